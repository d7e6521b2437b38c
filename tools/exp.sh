timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -5
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1_final_default.json 2> gpurun_out/bench_r1_final_default.err
python -c "
import json
for f in ('gpurun_out/bench_r1_final_default.json',):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), round(d['e2e']['value']), d['steps'], d['stage_ms'], d['roofline']['frac'], d['single_map_latency']); print(d['scan_front_end'])
"
tail -3 gpurun_out/bench_r1_final_default.err

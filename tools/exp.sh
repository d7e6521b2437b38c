timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -5
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1z.csv $B > gpurun_out/ncu_z.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r1_final_default.json 2> gpurun_out/bench_r1_final_default.err
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/bench_r1_final_steps8.json 2>> gpurun_out/bench_r1_final_default.err
python -c "
import json
for f in ('gpurun_out/bench_r1_final_default.json','gpurun_out/bench_r1_final_steps8.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), round(d['e2e']['value']), d['steps'], d['stage_ms'], d['roofline']['frac'], d['single_map_latency']['p50_ms']); print(d['scan_front_end'])
"
tail -3 gpurun_out/bench_r1_final_default.err

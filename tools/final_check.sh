timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1_final_default.json 2> gpurun_out/bench_r1_final_default.err; tail -c 300 gpurun_out/bench_r1_final_default.json
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/bench_r1_final_steps8.json 2>> gpurun_out/bench_r1_final_default.err; python -c "
import json
for f in ('gpurun_out/bench_r1_final_default.json','gpurun_out/bench_r1_final_steps8.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), round(d['e2e']['value']), d['steps'], d['stage_ms'], d['roofline']['frac'])
"

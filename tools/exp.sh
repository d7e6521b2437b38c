timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],1), d['stage_ms'], d['e2e']['value'], d['roofline']['frac'])"
timeout 200 python tools/run_batch.py 5000 1 16384

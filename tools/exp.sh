timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -3
for n in 1 8 32 150; do timeout 100 python tools/gpu_probe.py $n 2>&1 | grep -v "Mcycles\|^   " | tail -1; done
echo "--- batch 256 default"; timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],1), d['stage_ms'], round(d['e2e']['value']))"

timeout 300 python -m pytest tests/test_gpu_lsd.py -x -q 2>&1 | tail -2
LSDB_GROW_WARPS=4 timeout 60 python tools/gpu_sweep.py child
for nb in 1 3 4; do echo "--- inflight $nb"; timeout 400 python bench.py --steps $((nb*2)) --warmup 3 --no-cpu-baseline --inflight $nb 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],1), d['e2e']['value'])"; done

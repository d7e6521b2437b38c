timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -3
for nw in 4 16; do LSDB_GROW_WARPS=$nw timeout 60 python tools/gpu_sweep.py child; done
for nb in 1 4; do echo "--- inflight=$nb"; timeout 400 python bench.py --steps $((nb*2)) --warmup 3 --no-cpu-baseline --inflight $nb 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],1), round(d['e2e']['value']), d['stage_ms'])"; done

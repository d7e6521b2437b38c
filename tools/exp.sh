timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -4
for nw in 4 8; do echo "--- steal=0 nw=$nw single"; LSDB_STEAL=0 LSDB_GROW_WARPS=$nw timeout 60 python tools/gpu_sweep.py child; done
for nw in 4 8; do echo "--- batch 256 steal=0 nw=$nw"; LSDB_STEAL=0 LSDB_GROW_WARPS=$nw timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],1), d['stage_ms'], round(d['e2e']['value']))"; done

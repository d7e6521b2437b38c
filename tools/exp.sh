timeout 600 python -m pytest tests/test_gpu_fa.py tests/test_gpu_dropin.py tests/test_gpu_lsd.py -x -q -k "fa or dropin or scan" 2>&1 | grep -v "^Score" | tail -3
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --inflight 1 --maps-per-gpu 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['association'])"

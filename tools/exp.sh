timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -3
timeout 100 python tools/gpu_probe.py 32 2>&1 | grep -v "Mcycles\|^   " | tail -1

timeout 300 python tools/lat_probe.py 2>&1 | tail -24
timeout 300 python -m pytest tests/test_gpu_lsd.py -x -q 2>&1 | tail -3

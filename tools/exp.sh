timeout 300 python tools/scan_lsd_probe.py 4096 2>&1 | tail -6

timeout 300 python tools/scan_lsd_probe.py 512 2>&1 | tail -5
timeout 300 python tools/scan_lsd_probe.py 1024 2>&1 | tail -5
timeout 300 python tools/scan_lsd_probe.py 4096 2>&1 | tail -5

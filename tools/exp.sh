for nw in 4 16; do echo "--- nw=$nw single"; LSDB_GROW_WARPS=$nw timeout 60 python tools/gpu_sweep.py child; done
timeout 300 python tools/tail_probe.py 256 2>&1 | head -3

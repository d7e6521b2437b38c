#!/usr/bin/env python3
"""stage times of one batch of synthetic maps: tools/stage_probe.py [n_maps] [size] (development)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import synth
lsdb = load_package(); ctx = lsdb.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
maps = [synth.occupancy_grid(size, size, seed=1000 + (i % 16)) for i in range(n)]
b = lsdb.Batch(ctx, [(size, size)] * n, max_lines=16384 if size > 8192 else 4096); b.upload(maps)
for _ in range(3):
    b.run(); b.sync()
print("n", n, "size", size, "stage", b.stage_ms(), "segments", int(b.counts().sum()))

#!/usr/bin/env python3
"""run a batch of synthetic maps once: tools/run_batch.py first_seed count [size]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import synth
lsdb = load_package(); ctx = lsdb.Context(0)
first, cnt = int(sys.argv[1]), int(sys.argv[2]); size = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
maps = [synth.occupancy_grid(size, size, seed=first + i) for i in range(cnt)]
b = lsdb.Batch(ctx, [(size, size)] * cnt, max_lines=65536); b.upload(maps); b.run(); b.sync()
print("ok", b.counts().sum(), b.stage_ms())

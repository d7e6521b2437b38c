"""Times lsdb_feature_scan_frames on 10k lidar sweeps (the golden bundled frames, tiled) — GPU box only."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402

lsdb = load_package(); ctx = lsdb.Context(0)
g = np.load(os.path.join(ROOT, "tests", "golden", "lidar_frames.npz")); mp = g["map_param"]
fr = []
for f in range(int(g["n_frames"])):
    r, a = g[f"f{f}/ranges"], g[f"f{f}/angles"]; k = np.isfinite(r); fr.append((r[k], a[k]))
fr = (fr * 115)[:int(sys.argv[1]) if len(sys.argv) > 1 else 10000]
ctx.feature_scan(mp[2], mp[3], mp[4], fr[:50])
info, lines, loff, pts, poff = ctx.feature_scan(mp[2], mp[3], mp[4], fr, raw=True)
for _ in range(3):
    t = time.time(); ctx.feature_scan(mp[2], mp[3], mp[4], fr, raw=True, capacity=(len(lines), len(pts))); dt = time.time() - t
    print("frames", len(fr), "e2e s", round(dt, 4), "kernel ms (both kernels)", round(ctx.feature_scan_last_ms(), 3), "lines", len(lines), "pts", len(pts))

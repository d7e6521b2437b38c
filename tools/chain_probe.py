"""lidar sweeps -> pose estimates in one call (lsdb_scan_estimate_frames) on 10k sweeps — GPU box only; run under ncu for
the launch list of the chain."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402

lsdb = load_package(); ctx = lsdb.Context(0)
g = np.load(os.path.join(ROOT, "tests", "golden", "lidar_frames.npz")); mp = g["map_param"]
gf = np.load(os.path.join(ROOT, "tests", "golden", "fa_frames.npz")); gm = np.load(os.path.join(ROOT, "tests", "golden", "bundled_maps.npz"))
fr = []
for f in range(int(g["n_frames"])):
    r, a = g[f"f{f}/ranges"], g[f"f{f}/angles"]; k = np.isfinite(r); fr.append((r[k], a[k]))
fr = (fr * 115)[:10000]
fm = lsdb.FaMap(ctx, ctx.map_cache(gm["mapValue/map"], float(gm["mapValue/param"][2])), gf["map_lines"])
for _ in range(2):
    t = time.time(); info, est = fm.scan_estimate(mp[2], mp[3], mp[4], fr); dt = time.time() - t
    print("sweeps", len(fr), "s", round(dt, 4), "hypotheses", int(est["n_hyp"].sum()), "matched", int((est["n_kept"] > 0).sum()),
          "fscan ms", round(ctx.feature_scan_last_ms(), 3), "fa ms", round(fm.last_ms(), 3))

"""BASELINE configs[3] shape: lidar sweeps -> device FeatureScan rasters -> batched LSD on the rasters (GPU box only)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402

lsdb = load_package(); ctx = lsdb.Context(0)
g = np.load(os.path.join(ROOT, "tests", "golden", "lidar_frames.npz")); mp = g["map_param"]
base = []
for f in range(int(g["n_frames"])):
    r, a = g[f"f{f}/ranges"], g[f"f{f}/angles"]; k = np.isfinite(r); base.append((r[k], a[k]))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(7)
sweeps = [(base[i % len(base)][0] * (1 + rng.normal(0, 0.002, len(base[i % len(base)][0]))), base[i % len(base)][1]) for i in range(N)]
t = time.time(); fs = ctx.feature_scan(mp[2], mp[3], mp[4], sweeps, want_rasters=True); t_fs = time.time() - t
maps = [np.ascontiguousarray((o["line_im"] > 0).astype(np.uint8)) for o in fs]     # mapValue convention: occupied = 1
px = sum(m.size for m in maps)
for env in [dict(), dict(LSDB_GROW_WIDE="1"), dict(LSDB_GROW_WARPS="2", LSDB_GROW_WIDE="1"), dict(LSDB_SUPER_SHIFT="3"), dict(LSDB_NO_SMEM_BAN="1")]:
    os.environ.update(env)
    b = lsdb.Batch(ctx, [(m.shape[1], m.shape[0]) for m in maps], max_lines=256)
    for k in list(env): os.environ.pop(k)
    b.upload(maps); b.run(); b.sync()
    ts = []
    for _ in range(3):
        t = time.time(); b.run(); b.sync(); ts.append(time.time() - t)
    t = time.time(); b.upload(maps); b.run(); out = b.download(); t_e2e = time.time() - t
    print(f"{N} rasters {px/1e6:.0f} Mpx {str(env):30s} run {min(ts)*1e3:7.1f} ms = {N/min(ts):9.0f} rasters/s {px/min(ts)/1e6:8.0f} Mpx/s  e2e {t_e2e*1e3:.0f} ms  stages {b.stage_ms()}  segments {int(out['counts'].sum())}")
    b.close()
print("feature_scan with rasters (host out)", round(t_fs, 3), "s")

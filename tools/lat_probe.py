#!/usr/bin/env python3
"""single-map latency of the bundled maps under different team shapes (development probe)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
lsdb = load_package(); ctx = lsdb.Context(0)
g = np.load(os.path.join(ROOT, "tests", "golden", "bundled_maps.npz"))
for name in ("mapValue", "mapValue_aisle2", "mapValue_map1"):
    m = g[name + "/map"]
    for env in [dict(), dict(LSDB_SUPER_SHIFT="3"), dict(LSDB_SUPER_SHIFT="2"), dict(LSDB_SUPER_SHIFT="1"), dict(LSDB_SUPER_SHIFT="0"), dict(LSDB_SUPER_SHIFT="1", LSDB_GROW_WARPS="16"), dict(LSDB_SUPER_SHIFT="0", LSDB_GROW_WARPS="16")]:
        os.environ.update(env)
        b = lsdb.Batch(ctx, [(m.shape[1], m.shape[0])]); b.upload([m])
        for k in list(env): os.environ.pop(k)
        ts = []
        for _ in range(30):
            b.run(); b.sync(); ts.append(b.stage_ms()["grow"])
        st = b.stats()
        print(f"{name:16s} {str(env):60s} grow p50 {np.median(ts):6.2f} ms  respec={st['respec_evals']} grown_px={st['grown_px']} wait={st['cyc_wait']/1e6:.1f}M retire={st['cyc_retire']/1e6:.1f}M spec={st['cyc_spec']/1e6:.1f}M")
        b.close()

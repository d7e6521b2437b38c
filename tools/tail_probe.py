#!/usr/bin/env python3
"""per-map time distribution of a 256-map batch (development probe)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import synth
lsdb = load_package(); ctx = lsdb.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
maps = [synth.occupancy_grid(4096, 4096, seed=1000 + i) for i in range(n)]
b = lsdb.Batch(ctx, [(4096, 4096)] * n); b.upload(maps)
for _ in range(2):
    b.run(); b.sync()
st = [b.map_stats(i) for i in range(n)]
ms = np.array([s["ns_map"] / 1e6 for s in st]); cells = np.array([s["cells"] for s in st]); live = np.array([s["live_seeds"] for s in st])
gpx = np.array([s["grown_px"] for s in st]); spec = np.array([s["cyc_spec"] / 1e6 for s in st]); ret = np.array([s["cyc_retire"] / 1e6 for s in st])
print("stage", b.stage_ms())
print("ms per map: min %.0f p50 %.0f p90 %.0f max %.0f" % (ms.min(), np.median(ms), np.percentile(ms, 90), ms.max()))
o = np.argsort(-ms)[:8]
for i in o: print(f"  map {i}: {ms[i]:.0f} ms cells={cells[i]} live={live[i]} grown_px={gpx[i]} spec={spec[i]:.0f}M retire={ret[i]:.0f}M accepts={st[i]['accepts']} respec={st[i]['respec_evals']} cyc_respec={st[i]['cyc_respec']/1e6:.0f}M")
print("corr(ms, live)=%.2f corr(ms, grown_px)=%.2f corr(ms, cells)=%.2f" % (np.corrcoef(ms, live)[0, 1], np.corrcoef(ms, gpx)[0, 1], np.corrcoef(ms, cells)[0, 1]))
acc = np.array([s["accepts"] for s in st]); print("corr(ms, accepts)=%.2f  mean ms %.1f  sum/256 %.1f" % (np.corrcoef(ms, acc)[0, 1], ms.mean(), ms.sum() / n))
np.save(os.path.join(ROOT, "gpurun_out", "tail_ms.npy"), np.stack([ms, cells, live, gpx, acc]))

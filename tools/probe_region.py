#!/usr/bin/env python3
"""region-stage probe: tools/probe_region.py [n_maps] [size] -> stage times, all counters, per-map spread (development)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import synth
lsdb = load_package(); ctx = lsdb.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
maps = [synth.occupancy_grid(size, size, seed=1000 + i) for i in range(n)]
b = lsdb.Batch(ctx, [(size, size)] * n, max_lines=8192); b.upload(maps)
for _ in range(2):
    b.run(); b.sync()
print("env", {k: v for k, v in os.environ.items() if k.startswith("LSDB_")}, "n", n, "stage", b.stage_ms())
st = b.stats()
print({k: (round(v / 1e6) if k.startswith("cyc_") else v) for k, v in st.items()})
per = [b.map_stats(i) for i in range(n)]
ms = np.array([s["ns_map"] / 1e6 for s in per])
print("ms per map: min %.1f p50 %.1f p90 %.1f max %.1f" % (ms.min(), np.median(ms), np.percentile(ms, 90), ms.max()))
print("per map: large evals %.0f rounds %.0f respec %.0f (none %.0f conflict %.0f commit %.0f lost %.0f) requeue %.0f dropped %.0f big %.0f" % tuple(
    np.mean([f(s) for s in per]) for f in (lambda s: s["grows"] - s["small"], lambda s: s["rounds"], lambda s: s["respec_evals"], lambda s: s["rs_none"],
                                           lambda s: s["rs_conflict"], lambda s: s["rs_commit"], lambda s: s["rs_lost_commit"], lambda s: s["rs_requeue"],
                                           lambda s: s["rs_dropped"], lambda s: s["rs_big"])))
print("per map Mcycles: retire %.0f respec %.0f spec(worker) %.0f rounds %.0f wait %.0f | lane: grow %.0f rect %.0f nfa %.0f" % tuple(
    np.mean([s[k] for s in per]) / 1e6 for k in ("cyc_retire", "cyc_respec", "cyc_spec", "cyc_grow", "cyc_wait", "cyc_lane_grow", "cyc_rect", "cyc_nfa")))

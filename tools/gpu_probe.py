#!/usr/bin/env python3
"""Stage timings + pipeline counters on a GPU box (development probe, not a benchmark)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import synth
lsdb = load_package()
ctx = lsdb.Context(0)

def run(name, maps, reps=3):
    b = lsdb.Batch(ctx, [(m.shape[1], m.shape[0]) for m in maps])
    b.upload(maps)
    for _ in range(reps):
        t = time.time(); b.run(); b.sync(); dt = time.time() - t
    st = b.stats(); ms = b.stage_ms()
    px = sum(m.size for m in maps)
    print(f"{name}: n={len(maps)} wall={dt*1e3:.2f} ms stages={ {k: round(v,3) for k,v in ms.items()} } Mpx/s={px/dt/1e6:.1f}")
    print("   ", {k: st[k] for k in ("cells","live_seeds","grows","grown_px","nfa_calls","nfa_px","accepts","spec_evals","respec_evals","chunks")})
    print("    Mcycles:", {k: round(st[k]/1e6,2) for k in st if k.startswith("cyc_")})
    print("    per-map avg: %.2f ms, SM clock seen by clock64: %.0f MHz" % (st["ns_map"] / 1e6 / len(maps), st["cyc_map"] / max(1, st["ns_map"]) * 1e3))
    b.close()

g = np.load(os.path.join(ROOT, "tests", "golden", "bundled_maps.npz"))
names = ["mapValue", "mapValue_aisle1", "mapValue_aisle2", "mapValue_aisle3", "mapValue_map1", "mapValue_map2"]
run("config1 mapValue", [g["mapValue/map"]])
run("config2 bundled", [g[n + "/map"] for n in names])
sizes = [int(a) for a in sys.argv[1:]] or [1, 8]
for nb in sizes:
    maps = [synth.occupancy_grid(4096, 4096, seed=1000 + i) for i in range(nb)]
    run(f"4096^2 x{nb}", maps, reps=2)

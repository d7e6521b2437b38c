#!/usr/bin/env python3
"""createMapCache: device vs the oracle port on the host (development probe)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import synth, oraclebind
lsdb = load_package(); ctx = lsdb.Context(0)
for size in (1377, 4096):
    m = synth.occupancy_grid(size, size if size == 4096 else 428, seed=9) if size == 4096 else np.load(os.path.join(ROOT, "tests/golden/bundled_maps.npz"))["mapValue/map"]
    ctx.map_cache(m, 0.025)
    t = time.time(); a = ctx.map_cache(m, 0.025); dt = time.time() - t
    t = time.time(); b = oraclebind.map_cache(m, 0.025); dc = time.time() - t
    print(m.shape, "device %.1f ms (host buffers in/out)  cpu port %.1f ms  equal=%s" % (dt * 1e3, dc * 1e3, np.array_equal(a, b)))

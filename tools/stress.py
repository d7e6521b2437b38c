#!/usr/bin/env python3
"""Determinism stress: the same batch under several team shapes, several times; every run must give identical rectangles."""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import synth
lsdb = load_package(); ctx = lsdb.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
maps = [synth.occupancy_grid(4096, 4096, seed=3000 + i) for i in range(n)]
ref = None
for env in [dict(LSDB_GROW_WARPS="4"), dict(LSDB_GROW_WARPS="8"), dict(LSDB_GROW_WARPS="16"), dict(LSDB_GROW_WARPS="8", LSDB_STEAL="1"), dict(LSDB_GROW_WARPS="2")]:
    os.environ.update(env)
    b = lsdb.Batch(ctx, [(4096, 4096)] * n); b.upload(maps)
    for k in list(env): os.environ.pop(k)
    for rep in range(3):
        b.run()
        got = b.download(want_lines=False, want_rects=True)
        h = hashlib.sha1(b"".join(r.tobytes() for r in got["rects"]) + got["counts"].tobytes()).hexdigest()
        if ref is None: ref = h
        print(env, rep, h[:12], "OK" if h == ref else "MISMATCH", int(got["counts"].sum()), flush=True)
        assert h == ref
    b.close()
print("stress ok")

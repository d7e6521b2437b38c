#!/usr/bin/env python3
"""A/B of the two cuts of the stencil stage (stencil.cu: LSDB_STENCIL=1, stencil2.cu: default) on a GPU box (development).

    tools/stencil_ab.py dump OUT.npz        planes + segment tables of a fixed set of maps, stage times of a 64-map batch
    tools/stencil_ab.py compare A.npz B.npz bit-for-bit comparison with the location of the first differences
    tools/stencil_ab.py both REPORT.txt     the two dumps in child processes (one per version) and the comparison
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases():
    import synth
    out = []
    g = np.load(os.path.join(ROOT, "tests", "golden", "bundled_maps.npz"))
    for k in g.files:
        if k.endswith("/map"):
            out.append((k, np.ascontiguousarray(g[k])))
    out.append(("synth600x400", synth.occupancy_grid(600, 400, seed=7)))
    out.append(("synth1001x777", synth.occupancy_grid(1001, 777, seed=11)))
    out.append(("synth97x45", synth.occupancy_grid(97, 45, seed=12)))
    out.append(("synth40x40", synth.occupancy_grid(40, 40, seed=13)))
    out.append(("synth333x1500bw", synth.occupancy_grid(333, 1500, seed=14, border_walls=True)))
    for s in (1000, 1001, 1002):
        out.append((f"synth4096_{s}", synth.occupancy_grid(4096, 4096, seed=s)))
    out.append(("synth2048bw", synth.occupancy_grid(2048, 2048, seed=15, border_walls=True)))
    return out


def dump(path):
    from __graft_entry__ import load_package
    import synth
    lsdb = load_package()
    ctx = lsdb.Context(0)
    res = {}
    for name, m in cases():
        b = lsdb.Batch(ctx, [(m.shape[1], m.shape[0])])
        b.upload([m]); b.run()
        got = b.download(want_rects=True)
        pl = b.planes(0)
        res[name + "/mag"] = pl["mag"]; res[name + "/deg"] = pl["deg"]; res[name + "/used"] = pl["used"]
        res[name + "/labels"] = pl["labels"]; res[name + "/seeds"] = pl["seeds"]
        res[name + "/rects"] = np.asarray(got["rects"][0]); res[name + "/n"] = np.array([int(got["counts"][0])])
        res[name + "/launches"] = np.array([b.launches()])
        b.close()
    # timing: 64 maps of 4096^2, the stage alone
    n = 64
    maps = [synth.occupancy_grid(4096, 4096, seed=1000 + (i % 16)) for i in range(16)]
    maps = [maps[i % 16] for i in range(n)]
    b = lsdb.Batch(ctx, [(4096, 4096)] * n); b.upload(maps)
    st = []
    for _ in range(4):
        b.run(); b.sync(); st.append(b.stage_ms())
    res["timing"] = np.array([json.dumps(dict(n=n, stage_ms=st[-1], all=[s["stencil"] for s in st], segments=int(b.counts().sum())))])
    b.close(); ctx.close()
    np.savez(path, **res)
    print("dumped", path, res["timing"][0])


def compare(pa, pb, out=sys.stdout):
    A, B = np.load(pa), np.load(pb)
    bad = 0
    for k in A.files:
        if k == "timing" or k.endswith("/launches"):
            continue
        a, b = A[k], B[k]
        if a.shape != b.shape:
            print("DIFF", k, "shapes", a.shape, b.shape, file=out); bad += 1; continue
        same = np.array_equal(a.view(np.uint8), b.view(np.uint8)) if a.dtype.kind == "f" else np.array_equal(a, b)
        if not same:
            bad += 1
            av = a.view(np.int64) if a.dtype == np.float64 else a
            bv = b.view(np.int64) if b.dtype == np.float64 else b
            idx = np.argwhere(av != bv)
            print("DIFF", k, "count", len(idx), "of", a.size, "first", idx[:12].tolist(), file=out)
            for i in idx[:6]:
                t = tuple(i)
                print("    at", t, "tile-rel", [int(v) % 32 for v in t], "A", repr(a[t]), "B", repr(b[t]), file=out)
    print("timing A", A["timing"][0], file=out)
    print("timing B", B["timing"][0], file=out)
    print("launches A/B", [int(A[k][0]) for k in A.files if k.endswith("/launches")][:3], [int(B[k][0]) for k in B.files if k.endswith("/launches")][:3], file=out)
    print("RESULT", "IDENTICAL" if bad == 0 else f"{bad} arrays differ", file=out)
    return bad


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        dump(sys.argv[2])
    elif sys.argv[1] == "compare":
        sys.exit(1 if compare(sys.argv[2], sys.argv[3]) else 0)
    else:
        rep = sys.argv[2]
        e1 = dict(os.environ, LSDB_STENCIL="1")
        e2 = dict(os.environ); e2.pop("LSDB_STENCIL", None)
        variants = [("stencil2.cu, 1 tile per CTA", dict(e2, LSDB_STENCIL_G="1"), "/tmp/ab_g1.npz"),
                    ("stencil2.cu, 2 tiles per CTA", dict(e2, LSDB_STENCIL_G="2"), "/tmp/ab_g2.npz"),
                    ("stencil2.cu, 4 tiles per CTA", dict(e2, LSDB_STENCIL_G="4"), "/tmp/ab_g4.npz"),
                    ("stencil2.cu, 2 tiles per CTA, deferred pixels kept in the tile", dict(e2, LSDB_STENCIL_G="2", LSDB_STENCIL_DEFER="0"), "/tmp/ab_g2n.npz")]
        subprocess.check_call([sys.executable, __file__, "dump", "/tmp/ab_v1.npz"], env=e1)
        bad = 0
        with open(rep, "w") as f:
            for name, env, path in variants:
                subprocess.check_call([sys.executable, __file__, "dump", path], env=env)
                print("== stencil.cu (A) vs", name, "(B)", file=f)
                bad += compare("/tmp/ab_v1.npz", path, f)
        print(open(rep).read())
        sys.exit(1 if bad else 0)

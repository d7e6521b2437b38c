#!/usr/bin/env python3
"""One 4096^2 map through the region pipeline under different team sizes / run-ahead windows (development probe)."""
import os, sys, time, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from __graft_entry__ import load_package
    import synth
    lsdb = load_package(); ctx = lsdb.Context(0)
    size = int(os.environ.get("SWEEP_SIZE", "4096"))
    m = synth.occupancy_grid(size, size, seed=1000)
    b = lsdb.Batch(ctx, [(size, size)], max_lines=65536); b.upload([m])
    for _ in range(2):
        b.run(); b.sync()
    st = b.stats(); ms = b.stage_ms()
    keys = ("live_seeds", "grows", "grown_px", "small", "regrows", "nfa_calls", "rejects", "accepts", "spec_evals", "respec_evals", "rs_none", "rs_conflict", "rs_commit", "rs_lost_commit")
    print("grow %.1f ms | " % ms["grow"] + " ".join(f"{k}={st[k]}" for k in keys) + " | Mcyc " +
          " ".join(f"{k[4:]}={st[k]/1e6:.0f}" for k in st if k.startswith("cyc_")))
    sys.exit(0)
for nw, ra in [(1, 8), (4, 64), (16, 256)]:
    env = dict(os.environ, LSDB_GROW_WARPS=str(nw), LSDB_RUNAHEAD=str(ra))
    out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
    print(f"nw={nw:2d} runAhead={ra:3d}: {out.stdout.strip()} {out.stderr.strip()[-300:]}")

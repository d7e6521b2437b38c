timeout 300 python -m pytest tests/test_gpu_rdp.py -x -q 2>&1 | tail -15
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rdp.py -x -q -k "ragged or non_default or golden" 2>&1 | grep -v "^Score" | tail -12
timeout 200 python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, "tests")
from __graft_entry__ import load_package
lsdb = load_package(); ctx = lsdb.Context(0)
g = np.load("tests/golden/lidar_frames.npz"); mp = g["map_param"]
fr = []
for f in range(int(g["n_frames"])):
    r, a = g[f"f{f}/ranges"], g[f"f{f}/angles"]; k = np.isfinite(r); fr.append((r[k], a[k]))
fr = (fr * 115)[:10000]
ctx.feature_scan(mp[2], mp[3], mp[4], fr[:50])
info, lines, loff, pts, poff = ctx.feature_scan(mp[2], mp[3], mp[4], fr, raw=True)
for _ in range(3):
    t = time.time(); ctx.feature_scan(mp[2], mp[3], mp[4], fr, raw=True, capacity=(len(lines), len(pts))); dt = time.time() - t
    print("e2e s", round(dt, 4), "kernel ms (2 passes)", round(ctx.feature_scan_last_ms(), 3), "lines", len(lines), "pts", len(pts))
PY

"""Seeded synthetic inputs shaped like the reference's data.

occupancy_grid(): Karto-convention u8 grid — 0 free, 255 unknown (~20 %: outer margin + blobs), 1 occupied
(~1.2 %: room outlines and oblique wall polylines, 1-3 px thick), ~0.05 % speckle, and a few walls that run
into row 0 / column 0 to exercise the reference's border quirk (LSD/myLSD.cpp:135-136)."""
import numpy as np


def _draw_wall(m, x0, y0, x1, y1, thick, val=1):
    rows, cols = m.shape
    L = max(abs(x1 - x0), abs(y1 - y0))
    n = int(L * 2) + 2
    t = np.linspace(0.0, 1.0, n)
    xs = x0 + (x1 - x0) * t
    ys = y0 + (y1 - y0) * t
    d = np.hypot(x1 - x0, y1 - y0) + 1e-9
    nx, ny = -(y1 - y0) / d, (x1 - x0) / d
    for o in range(thick):
        xi = np.rint(xs + nx * o).astype(np.int64)
        yi = np.rint(ys + ny * o).astype(np.int64)
        ok = (xi >= 0) & (xi < cols) & (yi >= 0) & (yi < rows)
        m[yi[ok], xi[ok]] = val


def occupancy_grid(cols, rows, seed, occ_frac=0.012, speckle=0.0005, border_walls=False):
    rng = np.random.default_rng(seed)
    m = np.zeros((rows, cols), np.uint8)
    # unknown: outer margin of random width + a few rectangular blobs  (~20 %)
    mx0, mx1 = rng.integers(int(cols * 0.02), int(cols * 0.07) + 2, 2)
    my0, my1 = rng.integers(int(rows * 0.02), int(rows * 0.07) + 2, 2)
    m[:my0, :] = 255; m[rows - my1:, :] = 255; m[:, :mx0] = 255; m[:, cols - mx1:] = 255
    if not border_walls:
        # like every bundled map, the first row / column themselves are 0: the reference leaves them un-remapped,
        # so a 255 border line would be a ring-shaped edge that every one of its ~2(W'+H') pixels re-floods
        m[0, :] = 0; m[:, 0] = 0
    for _ in range(6):
        w, h = rng.integers(cols // 20, cols // 7 + 2), rng.integers(rows // 20, rows // 7 + 2)
        x, y = rng.integers(0, cols - w), rng.integers(0, rows - h)
        m[y:y + h, x:x + w] = 255
    target = int(occ_frac * rows * cols)
    scale = max(cols, rows)
    guard = 0
    occ = 0
    while occ < target and guard < 20000:
        guard += 1
        thick = int(rng.integers(1, 4))
        if rng.random() < 0.6:  # a room outline (axis-aligned, optionally slightly rotated)
            w, h = rng.integers(scale // 40 + 8, scale // 6 + 16), rng.integers(scale // 40 + 8, scale // 6 + 16)
            cx, cy = rng.integers(0, cols), rng.integers(0, rows)
            ang = 0.0 if rng.random() < 0.7 else rng.uniform(-0.5, 0.5)
            c, s = np.cos(ang), np.sin(ang)
            pts = [(-w / 2, -h / 2), (w / 2, -h / 2), (w / 2, h / 2), (-w / 2, h / 2)]
            pts = [(cx + c * a - s * b, cy + s * a + c * b) for a, b in pts]
            for k in range(4):
                if rng.random() < 0.85:
                    _draw_wall(m, pts[k][0], pts[k][1], pts[(k + 1) % 4][0], pts[(k + 1) % 4][1], thick)
                    occ += int(np.hypot(pts[k][0] - pts[(k + 1) % 4][0], pts[k][1] - pts[(k + 1) % 4][1])) * thick
        else:  # an oblique polyline
            x, y = rng.uniform(0, cols), rng.uniform(0, rows)
            ang = rng.uniform(0, 2 * np.pi)
            for _seg in range(int(rng.integers(1, 5))):
                L = rng.uniform(scale / 60 + 10, scale / 5 + 20)
                x2, y2 = x + L * np.cos(ang), y + L * np.sin(ang)
                _draw_wall(m, x, y, x2, y2, thick)
                occ += int(L) * thick
                x, y = x2, y2
                ang += rng.choice([-np.pi / 2, np.pi / 2, rng.uniform(-0.6, 0.6)])
    # walls running INTO the first row / column (the reference never remaps or thresholds row 0 / col 0,
    # LSD/myLSD.cpp:135-136,153-154, so regions may grow along the border).  `border_walls` additionally
    # lays walls ALONG row 0 / col 0: every one of their pixels then seeds a flood of the whole border ring
    # in the reference algorithm — a parity edge case (tests), not part of the throughput workload.
    _draw_wall(m, 0, rows * 0.3, cols * 0.2, rows * 0.3, 2)
    _draw_wall(m, cols * 0.4, 0, cols * 0.4, rows * 0.15, 2)
    if border_walls:
        _draw_wall(m, 0, 0, cols * 0.1, 0, 1)
        _draw_wall(m, 0, rows * 0.6, 0, rows * 0.8, 1)
    ns = int(speckle * rows * cols)
    m[rng.integers(0, rows, ns), rng.integers(0, cols, ns)] = 1
    return m


def fake_scan_frame(m, map_lines, seed, max_pts=1200):
    """A synthetic scan frame consistent with map `m`: occupied pixels around a random pose, moved into a scan
    frame by a rigid transform; scan lines = the map lines (len >= 40) moved the same way."""
    rng = np.random.default_rng(seed)
    rows, cols = m.shape
    ys, xs = np.nonzero(m == 1)
    c = rng.integers(0, len(xs))
    cx, cy = float(xs[c]), float(ys[c])
    d = np.hypot(xs - cx, ys - cy)
    sel = np.argsort(d)[:max_pts]
    th = rng.uniform(-np.pi, np.pi)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    P = np.stack([xs[sel] - cx, ys[sel] - cy], 1).astype(np.float64) @ R.T
    off = -P.min(0) + 5.0
    pts = np.floor(P + off)
    ml = np.asarray(map_lines, np.float64).reshape(-1, 10)
    near = [i for i in range(len(ml)) if ml[i, 8] >= 40 and np.hypot((ml[i, 4] + ml[i, 6]) / 2 - cx, (ml[i, 5] + ml[i, 7]) / 2 - cy) < 400][:12]
    if not near:
        near = [int(np.argmax(ml[:, 8]))] if len(ml) else []
    sl = np.zeros((len(near), 10))
    for j, i in enumerate(near):
        a = (np.array([ml[i, 4] - cx, ml[i, 5] - cy]) @ R.T) + off
        b = (np.array([ml[i, 6] - cx, ml[i, 7] - cy]) @ R.T) + off
        a, b = np.rint(a), np.rint(b)
        k = (b[1] - a[1]) / (b[0] - a[0]) if b[0] != a[0] else np.inf
        sl[j] = [k, 0, 0, 0, a[0], a[1], b[0], b[1], np.hypot(*(b - a)), 0]
    return dict(scan_lines=sl, pts=pts, lidar_pose=np.rint(off), last_pose=np.array([-1.0, -1.0, 0.0]))


def lidar_frame(seed, n_beams=360, dropout=0.12, noise=0.004):
    """A seeded synthetic lidar sweep (ranges, angles) in the format of the bundled Lidar.txt frames: the sensor sits in a
    randomly rotated rectangular room with a few interior wall pieces; beams are cast against the walls, jittered, and a
    fraction is dropped (the Inf beams the reference's readers discard, LSD/main_on_windows.cpp:110-123).  Angles follow
    the bundled convention -3.12414 + i * (2*pi/n_beams)."""
    rng = np.random.default_rng(seed)
    ang = -3.12414 + np.arange(n_beams) * (2 * np.pi / n_beams)
    phi = rng.uniform(0, np.pi)
    hw, hh = rng.uniform(1.5, 9.0), rng.uniform(1.0, 5.0)
    ox, oy = rng.uniform(-0.6, 0.6) * hw, rng.uniform(-0.6, 0.6) * hh
    c = np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) - [ox, oy]
    segs = [(c[i], c[(i + 1) % 4]) for i in range(4)]
    for _ in range(int(rng.integers(0, 5))):
        p = np.array([rng.uniform(-hw, hw) - ox, rng.uniform(-hh, hh) - oy])
        d = rng.uniform(0.4, 2.5) * (np.array([1.0, 0.0]) if rng.random() < 0.5 else np.array([0.0, 1.0]))
        segs.append((p, p + d))
    R = np.array([[np.cos(phi), -np.sin(phi)], [np.sin(phi), np.cos(phi)]])
    dx, dy = np.cos(ang), np.sin(ang)
    best = np.full(n_beams, np.inf)
    for a, b in segs:
        a, b = R @ a, R @ b
        ex, ey = b - a
        den = dx * ey - dy * ex
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (a[0] * ey - a[1] * ex) / den
            u = (a[0] * dy - a[1] * dx) / den
        hit = (np.abs(den) > 1e-12) & (t > 0.05) & (u >= 0) & (u <= 1)
        best = np.where(hit & (t < best), t, best)
    best = best + rng.normal(0, noise, n_beams)
    best[rng.random(n_beams) < dropout] = np.inf
    best[best > 16.4] = np.inf
    keep = np.isfinite(best)
    return best[keep].copy(), ang[keep].copy()

"""ctypes bindings for the compiled reference (oracle/_ref/libref_*.so) — test infrastructure.

The .so files wrap the UNMODIFIED reference (see oracle/ref_harness.cpp); they are built in the
authoring container by oracle/Makefile and travel to the GPU box as prebuilt files."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
LSD_PARAMS = dict(sca=0.3, sig=0.6, angThre=22.5, denThre=0.7, pseBin=1024)  # LSD/baseFunc.h:64-68

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")


def available(variant="glibc"):
    return os.path.exists(os.path.join(REF_DIR, f"libref_{variant}.so"))


_libs = {}


def lib(variant="glibc"):
    if variant in _libs:
        return _libs[variant]
    L = C.CDLL(os.path.join(REF_DIR, f"libref_{variant}.so"))
    L.ref_variant.restype = C.c_char_p
    L.ref_lsd.restype = C.c_int
    L.ref_lsd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                          C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ref_create_map_cache.restype = None
    L.ref_create_map_cache.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
    L.ref_fa_scores.restype = C.c_int
    L.ref_fa_scores.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.ref_feature_association.restype = None
    L.ref_feature_association.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_feature_scan.restype = C.c_int
    L.ref_feature_scan_many.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
    L.ref_feature_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    _libs[variant] = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ref_lsd(map_u8, variant="glibc", want_maps=True, **kw):
    """Run the reference myLineSegmentDetector; returns a dict of numpy arrays."""
    p = dict(LSD_PARAMS); p.update(kw)
    m = np.ascontiguousarray(map_u8, dtype=np.uint8)
    rows, cols = m.shape
    W, H = int(np.floor(cols * p["sca"])), int(np.floor(rows * p["sca"]))
    max_lines = 65536
    lines = np.zeros((max_lines, 10), np.float64)
    out = dict(line_im=np.zeros((rows, cols), np.uint8), map_out=np.zeros((rows, cols), np.uint8))
    if want_maps:
        out.update(gauss=np.zeros((H, W)), mag=np.zeros((H, W)), deg=np.zeros((H, W)),
                   used=np.zeros((H, W), np.uint8), reg_idx=np.zeros((H, W), np.uint8))
    max_seeds = H * W
    seeds = np.zeros((max_seeds, 3), np.int32)
    ns = C.c_int(0)
    g = out.get
    n = lib(variant).ref_lsd(_p(m), cols, rows, p["sca"], p["sig"], p["angThre"], p["denThre"], p["pseBin"],
                             _p(lines), max_lines, _p(out["line_im"]), _p(out["map_out"]), _p(g("gauss")),
                             _p(g("mag")), _p(g("deg")), _p(g("used")), _p(g("reg_idx")), _p(seeds), max_seeds,
                             C.byref(ns))
    out["lines"] = lines[:n].copy()
    out["seeds"] = seeds[:ns.value].copy()
    out["n"] = n
    return out


def ref_map_cache(map_u8, res, variant="glibc"):
    m = np.ascontiguousarray(map_u8, dtype=np.uint8)
    rows, cols = m.shape
    out = np.zeros((rows, cols), np.float64)
    lib(variant).ref_create_map_cache(_p(m), cols, rows, float(res), _p(out))
    return out


def ref_fa_scores(scan_lines, map_lines, pts, map_cache, lidar_pose, last_pose, variant="glibc"):
    sl = np.ascontiguousarray(scan_lines, np.float64).reshape(-1, 10)
    ml = np.ascontiguousarray(map_lines, np.float64).reshape(-1, 10)
    pt = np.ascontiguousarray(pts, np.float64).reshape(-1, 2)
    mc = np.ascontiguousarray(map_cache, np.float64)
    rows, cols = mc.shape
    lp = np.asarray(lidar_pose, np.float64); la = np.asarray(last_pose, np.float64)
    cap = max(4 * len(sl) * len(ml), 4)
    idx = np.zeros((cap, 3), np.int32); val = np.zeros((cap, 4), np.float64)
    n = lib(variant).ref_fa_scores(_p(sl), len(sl), _p(ml), len(ml), _p(pt), len(pt), _p(mc), cols, rows,
                                   _p(lp), _p(la), _p(idx), _p(val), cap)
    return idx[:n].copy(), val[:n].copy()


def ref_feature_scan(map_param, ranges, angles, variant="glibc"):
    mp = np.asarray(map_param, np.float64)
    r = np.ascontiguousarray(ranges, np.float64); a = np.ascontiguousarray(angles, np.float64)
    lines = np.zeros((360, 10)); pts = np.zeros((200000, 2)); npts = C.c_int(0)
    lidar = np.zeros(2); imsz = np.zeros(2, np.int32)
    cap = 4096 * 4096
    im = np.zeros(cap, np.uint8)
    n = lib(variant).ref_feature_scan(_p(mp), _p(r), _p(a), len(r), _p(lines), 360, _p(pts), len(pts),
                                      C.byref(npts), _p(lidar), _p(imsz), _p(im), cap)
    w, h = int(imsz[0]), int(imsz[1])
    return dict(lines=lines[:n].copy(), pts=pts[:npts.value].copy(), lidar_pos=lidar, size=(w, h),
                line_im=im[:w * h].reshape(h, w).copy())


def ref_feature_scan_many(map_param, frames, variant="glibc"):
    """myrdp::FeatureScan over a list of (ranges, angles) frames inside one native call; returns (lines, samples)."""
    mp = np.asarray(map_param, np.float64)
    boff = np.zeros(len(frames) + 1, np.int32)
    boff[1:] = np.cumsum([len(r) for r, _ in frames])
    r = np.ascontiguousarray(np.concatenate([f[0] for f in frames]), np.float64)
    a = np.ascontiguousarray(np.concatenate([f[1] for f in frames]), np.float64)
    L = lib(variant)
    L.ref_feature_scan_many.restype = C.c_longlong
    npts = C.c_longlong(0)
    nl = L.ref_feature_scan_many(_p(mp), _p(r), _p(a), _p(boff), len(frames), C.byref(npts))
    return int(nl), int(npts.value)

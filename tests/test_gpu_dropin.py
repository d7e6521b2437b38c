"""The C++ drop-in boundary: oracle/_ref/libref_dropin.so links the reference's OWN host code (the RDP helpers,
ukf, the harness that mirrors main_on_windows.cpp) with this repo's bodies for mylsd::myLineSegmentDetector, createMapCache, myrdp::FeatureScan and
myfa::FeatureAssociation (linesegmentdetector-slam_b200/host/*.cpp -> liblsdb200.so).  Same entry points, same structs:
the outputs must match the unmodified reference (libref_glibc.so / golden fixtures)."""
import os

import numpy as np
import pytest

import refbind

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["mapValue", "mapValue_aisle1", "mapValue_aisle2", "mapValue_aisle3", "mapValue_map1", "mapValue_map2"]

needs_dropin = pytest.mark.skipif(not refbind.available("dropin"), reason="oracle/_ref/libref_dropin.so not built")


@needs_dropin
def test_myLineSegmentDetector_dropin_matches_reference_outputs():
    assert refbind.lib("dropin").ref_variant() == b"dropin"
    g = np.load(os.path.join(GOLD, "bundled_maps.npz"))
    for n in NAMES:
        m = g[n + "/map"]
        r = refbind.ref_lsd(m, variant="dropin", want_maps=False)
        assert r["n"] == len(g[n + "/lines"])
        assert np.array_equal(r["lines"], g[n + "/lines"], equal_nan=True)          # structLinesInfo table, bit for bit
        assert np.array_equal(np.packbits(r["line_im"] > 0), g[n + "/line_im_bits"])  # structLSD::lineIm
        if refbind.available("glibc"):
            ref = refbind.ref_lsd(m, variant="glibc", want_maps=False)
            assert np.array_equal(r["map_out"], ref["map_out"])                      # the in-place remap of the caller's Mat


@needs_dropin
def test_FeatureAssociation_dropin_matches_reference():
    """Whole FeatureAssociation behind the reference's entry point: device scoring + the host HMM gate / weighted mean of
    host/myFA_b200.cpp + the reference's own ukf.  The shipped reference drops queued tasks when it tears its pool down
    (LSD/myFA.cpp:61-63), so its output varies run to run; the expectation is therefore built from the reference's OWN
    serial scoring functions (golden fixture: NormalizedLineDirection / rotateScanIm / CalcScore per hypothesis), the
    selection rules of LSD/myFA.cpp:65-171 restated here, and the reference's ukf (libref_glibc.so: ref_ukf)."""
    if not refbind.available("glibc"):
        pytest.skip("reference library not built")
    import ctypes as C
    g = np.load(os.path.join(GOLD, "fa_frames.npz"))
    gm = np.load(os.path.join(GOLD, "bundled_maps.npz"))
    res = float(gm["mapValue/param"][2])
    mc = refbind.ref_map_cache(gm["mapValue/map"], res, variant="dropin")   # mylsd::createMapCache -> lsdb_map_cache (device)
    assert np.array_equal(mc, refbind.ref_map_cache(gm["mapValue/map"], res, variant="glibc"))
    rows, cols = mc.shape
    ml = np.ascontiguousarray(g["map_lines"], np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L = refbind.lib("glibc")
    L.ref_ukf.restype = None
    L.ref_ukf.argtypes = [C.c_void_p] * 4
    scan_pose = np.array([0.1, -0.05, 0.02])
    P0 = np.diag([100, 100, 100, 1, 1, 1, .1, .1, .1]).astype(np.float64)

    def dropin(f, last):
        sl = np.ascontiguousarray(g[f"f{f}/scan_lines"], np.float64); pts = np.ascontiguousarray(g[f"f{f}/pts"], np.float64)
        lid = np.ascontiguousarray(g[f"f{f}/lidar_pose"], np.float64); last = np.ascontiguousarray(last, np.float64)
        kx = np.array([last[0], last[1], last[2], 0, 0, 0, 0, 0, 0], np.float64); kP = np.ascontiguousarray(P0.copy())
        refbind.lib("dropin").ref_feature_association(p(sl), len(sl), p(ml), len(ml), p(pts), len(pts), p(mc), cols, rows,
                                                      p(lid), p(last), p(scan_pose), p(kx), p(kP))
        return kx, kP

    tracked = 0
    for f in range(int(g["n_frames"])):
        lid = g[f"f{f}/lidar_pose"]
        # first frame of a chain: the best hypothesis becomes the pose (LSD/myFA.cpp:100-110)
        idx, val = refbind.ref_fa_scores(g[f"f{f}/scan_lines"], ml, g[f"f{f}/pts"], mc, lid, [-1, -1, 0])
        keep = val[:, 3] < 3
        kx, kP = dropin(f, [-1.0, -1.0, 0.0])
        if not keep.any():
            assert np.array_equal(kx, [-1, -1, 0, 0, 0, 0, 0, 0, 0]) and np.array_equal(kP, P0)
            continue
        best = val[keep][np.argmin(val[keep, 3])]
        assert np.allclose(kx[:3], best[:3], rtol=1e-12, atol=1e-9), (f, kx[:3], best)
        assert np.array_equal(kx[3:], np.zeros(6)) and np.array_equal(kP, P0)
        # tracking: gate 60 px around the last pose, 1/score^2 weighted mean (:160-171), then the reference's ukf
        last = np.array([best[0] + 4.0, best[1] - 3.0, best[2] + 0.5])
        idx, val = refbind.ref_fa_scores(g[f"f{f}/scan_lines"], ml, g[f"f{f}/pts"], mc, lid, last)
        keep = val[:, 3] < 3
        kx, kP = dropin(f, last)
        if not keep.any():
            assert np.array_equal(kx, [-1, -1, 0, 0, 0, 0, 0, 0, 0])
            continue
        o = np.argsort(val[keep, 3], kind="stable")
        v = val[keep][o]
        w = 1.0 / v[:, 3] ** 2
        est = np.array([np.sum(v[:, 0] * w), np.sum(v[:, 1] * w), np.sum(v[:, 2] * w)]) / np.sum(w)
        ex = np.array([last[0], last[1], last[2], 0, 0, 0, 0, 0, 0], np.float64); eP = np.ascontiguousarray(P0.copy())
        L.ref_ukf(p(ex), p(eP), p(scan_pose), p(np.ascontiguousarray(est)))
        assert np.allclose(kx, ex, rtol=1e-9, atol=1e-9), (f, kx, ex)
        assert np.allclose(kP, eP, rtol=1e-9, atol=1e-9)
        tracked += 1
    assert tracked >= 6


@needs_dropin
def test_FeatureScan_dropin_matches_reference():
    """myrdp::FeatureScan through the reference's own header and structs (structFeatureScan: lineIm Mat, malloc'd
    linesInfo, lidarPos, scanImPoint vector), body = liblsdb200"""
    g = np.load(os.path.join(GOLD, "lidar_frames.npz"))
    mp = list(g["map_param"])
    for f in range(0, int(g["n_frames"]), 5):
        r, a = g[f"f{f}/ranges"], g[f"f{f}/angles"]
        keep = np.isfinite(r)
        fs = refbind.ref_feature_scan(mp, r[keep], a[keep], variant="dropin")
        assert np.array_equal(fs["lines"], g[f"f{f}/lines"], equal_nan=True)
        assert np.array_equal(fs["pts"], g[f"f{f}/pts"].astype(np.float64))
        assert np.array_equal(fs["lidar_pos"], g[f"f{f}/lidar_pos"]) and tuple(fs["size"]) == tuple(g[f"f{f}/size"])
        im = np.zeros(fs["line_im"].shape, np.uint8)
        im[g[f"f{f}/pts"][:, 1], g[f"f{f}/pts"][:, 0]] = 255
        assert np.array_equal(fs["line_im"], im)

"""SURVEY.md §8 f4, association half: the catkin snapshot's 13-argument myfa::FeatureAssociation (ROS/lsd/include/FeatureAssociation.h:46-60,
ROS/lsd/src/FeatureAssociation.cpp:36-299 — length filter, four pairings, RotateScanIm, the ray re-projection score).

not gpu : the plain-C restatement (oracle/fa_legacy_oracle.c) against golden vectors made by the UNMODIFIED source
          (tests/golden/fa_legacy.npz, tools/make_golden_fa_legacy.py) and, where oracle/_ref is built, bit for bit against the
          unmodified source with lsd_math.h bound in (libref_rosfa_lsdm.so);
gpu     : lsdb_fa_legacy through the C ABI — bit-exact against the restatement (scores included: the distances are added in ray
          order), the goldens, and the drop-in C++ body behind the reference's own signature (libdropin_rosfa.so)."""
import ctypes as C
import os

import numpy as np
import pytest

import oraclebind
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF_LSDM = os.path.join(ROOT, "oracle", "_ref", "libref_rosfa_lsdm.so")
DROPIN = os.path.join(ROOT, "oracle", "_ref", "libdropin_rosfa.so")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "fa_legacy.npz"))


@pytest.fixture(scope="module")
def cache():
    g = np.load(os.path.join(GOLD, "bundled_maps.npz"))
    mc = oraclebind.map_cache(g["mapValue/map"], float(g["mapValue/param"][2]))
    mc[mc == 1.0] = 2.0        # unreached cells as the snapshot's three-argument createMapCache leaves them (z_occ_max_dis = 2)
    return mc


def _frame(gold, f):
    return gold[f"f{f}/scan_lines"], gold[f"f{f}/lidar_pos"], gold[f"f{f}/ranges"], gold[f"f{f}/angles"]


def _check_vs_gold(gold, f, pose, est, real, exact_scores):
    want = gold[f"f{f}/pose_all"]
    assert pose.shape == want.shape
    if len(want) == 0:
        return
    cols = [0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
    assert np.array_equal(pose[:, cols], want[:, cols])                         # poses, end points, indices: same bits as the reference
    fin = np.isfinite(want[:, 3])
    assert np.array_equal(np.isfinite(pose[:, 3]), fin) and np.array_equal(np.isnan(pose[:, 3]), np.isnan(want[:, 3]))
    if exact_scores:
        assert np.array_equal(pose[:, 3], want[:, 3], equal_nan=True)
    else:
        assert np.allclose(pose[fin, 3], want[fin, 3], rtol=1e-12, atol=0)
    assert np.array_equal(est, gold[f"f{f}/est"]) and np.array_equal(real, gold[f"f{f}/est_real"])


def test_oracle_vs_golden_vectors_of_the_unmodified_source(gold, cache):
    some = 0
    for f in range(int(gold["n_frames"])):
        sl, lp, r, a = _frame(gold, f)
        pose, est, real = oraclebind.fa_legacy(sl, gold["map_lines"], float(gold["resol"]), gold["ori"], lp, cache, r, a)
        _check_vs_gold(gold, f, pose, est, real, exact_scores=True)
        some += int(np.isfinite(pose[:, 3]).sum()) if len(pose) else 0
    assert some > 500


@pytest.mark.skipif(not os.path.exists(REF_LSDM), reason="oracle/_ref/libref_rosfa_lsdm.so not built")
def test_oracle_equals_the_unmodified_source_bit_for_bit(gold, cache):
    R = C.CDLL(REF_LSDM); R.ros_fa.restype = C.c_int
    g = np.load(os.path.join(GOLD, "bundled_maps.npz"))
    m = g["mapValue/map"]
    for seed in range(12):
        fr = synth.fake_scan_frame(m, gold["map_lines"], seed=300 + seed)
        r, a = synth.lidar_frame(400 + seed, n_beams=(360, 720, 1081)[seed % 3])
        args = (fr["scan_lines"], gold["map_lines"], 0.05 if seed % 4 == 3 else float(gold["resol"]), (-3.5, 2.25), fr["lidar_pose"], cache, r, a)
        o = oraclebind.fa_legacy(*args)
        w = oraclebind.fa_legacy(*args, fn=R.ros_fa)
        assert o[0].shape == w[0].shape and len(o[0]) > 0
        assert np.array_equal(o[0], w[0], equal_nan=True) and np.array_equal(o[1], w[1]) and np.array_equal(o[2], w[2])


@pytest.mark.gpu
def test_cuda_legacy_association_vs_oracle_and_goldens(lsdb, ctx, gold, cache):
    fm = lsdb.FaMap(ctx, cache, gold["map_lines"])
    for f in range(int(gold["n_frames"])):
        sl, lp, r, a = _frame(gold, f)
        pose, est, real = fm.legacy(sl, float(gold["resol"]), gold["ori"], lp, r, a)
        o = oraclebind.fa_legacy(sl, gold["map_lines"], float(gold["resol"]), gold["ori"], lp, cache, r, a)
        assert pose.shape == o[0].shape
        if len(pose) == 0:
            assert est is None and real is None
            continue
        assert np.array_equal(pose, o[0], equal_nan=True)                       # every column, scores included: same bits as the oracle
        assert np.array_equal(est, o[1]) and np.array_equal(real, o[2])
        _check_vs_gold(gold, f, pose, est, real, exact_scores=True)
    assert fm.last_ms() > 0
    # seeded frames beyond the goldens, other resolutions (the length filter and the ray grid scale with it)
    g = np.load(os.path.join(GOLD, "bundled_maps.npz"))
    for seed in range(10):
        fr = synth.fake_scan_frame(g["mapValue/map"], gold["map_lines"], seed=500 + seed)
        r, a = synth.lidar_frame(600 + seed, n_beams=(360, 1081)[seed % 2])
        res = (0.025, 0.05, 0.1)[seed % 3]
        pose, est, real = fm.legacy(fr["scan_lines"], res, (1.5, -2.0), fr["lidar_pose"], r, a)
        o = oraclebind.fa_legacy(fr["scan_lines"], gold["map_lines"], res, (1.5, -2.0), fr["lidar_pose"], cache, r, a)
        assert np.array_equal(pose, o[0], equal_nan=True) and np.array_equal(est, o[1]) and np.array_equal(real, o[2])
    # a table that is too small is an error, not a silent truncation
    sl, lp, r, a = _frame(gold, 0)
    with pytest.raises(lsdb.LsdbError, match="CAPACITY"):
        fm.legacy(sl, float(gold["resol"]), gold["ori"], lp, r, a, max_cols=8)
    fm.close()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/libdropin_rosfa.so not built")
def test_dropin_body_behind_the_snapshots_signature(lsdb, ctx, gold, cache):
    D = C.CDLL(DROPIN); D.ros_fa.restype = C.c_int
    for f in range(int(gold["n_frames"])):
        sl, lp, r, a = _frame(gold, f)
        pose, est, real = oraclebind.fa_legacy(sl, gold["map_lines"], float(gold["resol"]), gold["ori"], lp, cache, r, a, fn=D.ros_fa)
        if len(gold[f"f{f}/pose_all"]) == 0:
            assert len(pose) == 0
            continue
        _check_vs_gold(gold, f, pose, est, real, exact_scores=True)

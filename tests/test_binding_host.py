"""Host-side marshalling of the ctypes binding (no GPU): record layouts and the ragged-sweep packing the C ABI expects."""
import numpy as np


def test_line_records_round_trip(lsdb):
    rng = np.random.default_rng(3)
    a = rng.normal(size=(7, 10)); a[:, 9] = rng.choice([-1, 1], 7)
    rec = lsdb.array_to_lines(a)
    assert rec.dtype == lsdb.LINE_DTYPE and rec.dtype.itemsize == 80
    assert np.array_equal(lsdb.lines_to_array(rec), a)


def test_sweep_packing(lsdb):
    sweeps = [(np.array([1.0, 2.0, 3.0]), np.array([0.1, 0.2, 0.3])), (np.array([4.0]), np.array([0.4])),
              (np.arange(5, dtype=np.float32), np.arange(5) * 0.5)]
    nf, boff, rng, ang, prm = lsdb._marshal_sweeps(sweeps, dict(thre_line=0.1))
    assert nf == 3 and boff.dtype == np.int32 and list(boff) == [0, 3, 4, 9]
    assert rng.dtype == np.float64 and rng.flags["C_CONTIGUOUS"] and list(rng[:4]) == [1.0, 2.0, 3.0, 4.0] and len(ang) == 9
    assert (prm.least_point, prm.thre_line, prm.least_dist_m) == (3, 0.1, 0.5)          # defaults of LSD/baseFunc.h:70-72
    nf, boff, rng, ang, _ = lsdb._marshal_sweeps([], {})
    assert nf == 0 and list(boff) == [0] and len(rng) == 0


def test_scan_record_layouts(lsdb):
    assert lsdb.SCAN_INFO_DTYPE.itemsize == 32 and lsdb.EST_DTYPE.itemsize == 72
    assert [lsdb.SCAN_INFO_DTYPE.fields[f][1] for f in ("n_lines", "n_pts", "im_cols", "im_rows", "lidar_x", "lidar_y")] == [0, 4, 8, 12, 16, 24]

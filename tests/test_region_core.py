"""The lane-serial core of the region stage (csrc/region_core.h), compiled for the host and driven by the reference's
sequential seed loop, against the oracle: used-map, labels, rectangles, log-NFA and the work counters bit for bit.
Pins the arithmetic / control flow every GPU lane executes (the concurrent part is pinned by the -m gpu tests)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oraclebind
import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "linesegmentdetector-slam_b200", "csrc")
SO = os.path.join(HERE, "_region_core.so")


@pytest.fixture(scope="module")
def core():
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-I/usr/local/cuda/include", "-I" + CSRC,
                           "-shared", "-o", SO, os.path.join(HERE, "region_core_host.cpp")])
    L = C.CDLL(SO)
    L.rgcore_lsd.restype = C.c_int
    L.rgcore_lsd.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return L


def _run(core, m, lane_cap=0, scout=0, **kw):
    o = oraclebind.lsd(m, want_line_im=False, **kw)
    p = dict(oraclebind.LSD_PARAMS); p.update(kw)
    H, W = o["mag"].shape
    seeds = np.ascontiguousarray(o["seeds"][:, 2] * W + o["seeds"][:, 1], np.int32)
    used = np.zeros((H, W), np.uint8); labels = np.zeros((H, W), np.int32); rects = np.zeros((65536, 13)); st = np.zeros(10, np.int64)
    n = core.rgcore_lsd(W, H, o["mag"].ctypes.data, o["deg"].ctypes.data, seeds.ctypes.data, len(seeds), p["sca"], p["angThre"], p["denThre"],
                        lane_cap, scout, used.ctypes.data, labels.ctypes.data, rects.ctypes.data, len(rects), st.ctypes.data)
    return o, n, used, labels, rects, st


def _check(core, m, **kw):
    _check1(core, m, scout=1, **kw)
    return _check1(core, m, scout=0, **kw)


def _check1(core, m, **kw):
    o, n, used, labels, rects, st = _run(core, m, **kw)
    assert n == o["n"]
    assert np.array_equal(used, o["used"]) and np.array_equal(labels, o["labels"])
    assert np.array_equal(rects[:n], o["rects"], equal_nan=True)
    s = o["stats"]
    got = dict(zip("live_seeds grows grown_px small regrows rrr_passes nfa_calls nfa_px rejects accepts".split(), st.tolist()))
    for k, v in got.items():
        assert v == s[k], (k, v, s[k])
    return o


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "bundled_maps.npz"))


@pytest.mark.parametrize("name", ["mapValue", "mapValue_aisle1", "mapValue_aisle2", "mapValue_aisle3", "mapValue_map1"])
def test_lane_core_on_bundled_maps(core, gold, name):
    _check(core, gold[name + "/map"])


@pytest.mark.parametrize("shape,seed,bw", [((600, 400), 7, False), ((333, 901), 22, True), ((64, 50), 23, False), ((1500, 1100), 31, True)])
def test_lane_core_on_synthetic_maps(core, shape, seed, bw):
    _check(core, synth.occupancy_grid(shape[0], shape[1], seed, border_walls=bw))


def test_lane_core_other_parameters(core, gold):
    _check(core, gold["mapValue_aisle2/map"], angThre=30.0, denThre=0.8)
    _check(core, gold["mapValue_aisle1/map"], angThre=15.0, denThre=0.6)


def test_lane_core_4096(core):
    _check(core, synth.occupancy_grid(4096, 4096, 1003))


def test_lane_overflow_is_reported(core, gold):
    o, n, *_ = _run(core, gold["mapValue_aisle1/map"], lane_cap=24)
    assert n == -1   # a region outgrew the lane's list: RG_OC_DEFER

"""K7's tracker core (csrc/k7_bytetrack_core.h) compiled for the host (tests/native, TEST BUILD ONLY: the product has
no CPU path) against scipy's linear_sum_assignment, supervision's linear_assignment as restated in the oracle, and the
restated ByteTrack itself on synthetic clips — ids and kept detections must be identical frame by frame."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from oracle.bytetrack_restated import ByteTrack as RefByteTrack, linear_assignment
from test_tracker import INIT, MAIN, synthetic_detections

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    spec = importlib.util.spec_from_file_location("bt_native_build", os.path.join(HERE, "native", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    L = C.CDLL(mod.build())
    L.bt_host_create.restype = C.c_void_p
    L.bt_host_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    L.bt_host_destroy.argtypes = [C.c_void_p]
    L.bt_host_reset.argtypes = [C.c_void_p]
    L.bt_host_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p]
    L.bt_host_assign.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bt_host_lsap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return L


class HostCore:
    def __init__(self, lib, track_activation_threshold=0.25, lost_track_buffer=30, minimum_matching_threshold=0.8,
                 frame_rate=30, minimum_consecutive_frames=1):
        self.lib = lib
        self.h = lib.bt_host_create(track_activation_threshold, track_activation_threshold + 0.1, minimum_matching_threshold,
                                    int(frame_rate / 30.0 * lost_track_buffer), minimum_consecutive_frames)

    def __del__(self):
        self.lib.bt_host_destroy(self.h)

    def update(self, xyxy, conf, cls=None, min_conf=-np.inf, class_mask=0xFFFFFFFF):
        xyxy = np.ascontiguousarray(xyxy, np.float32).reshape(-1, 4)
        conf = np.ascontiguousarray(conf, np.float32)
        cls = np.ascontiguousarray(cls, np.int32) if cls is not None else None
        n = len(conf)
        row, tid = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        k = self.lib.bt_host_update(self.h, xyxy.ctypes.data, conf.ctypes.data, cls.ctypes.data if cls is not None else None, n,
                                    float(min_conf), class_mask, row.ctypes.data, tid.ctypes.data)
        assert k >= 0
        return row[:k].astype(int), tid[:k].astype(int)


def test_lsap_equals_scipy_including_ties(lib):
    rng = np.random.default_rng(0)
    for trial in range(400):
        nr, nc = int(rng.integers(1, 14)), int(rng.integers(1, 14))
        if nr > nc:
            nr, nc = nc, nr
        kind = trial % 4
        if kind == 0:
            cost = rng.random((nr, nc))
        elif kind == 1:
            cost = rng.integers(0, 3, (nr, nc)).astype(np.float64)           # heavy ties
        elif kind == 2:
            cost = np.full((nr, nc), 0.8001)                                 # constant: scipy returns the identity
        else:
            cost = np.where(rng.random((nr, nc)) < 0.6, 0.8001, rng.random((nr, nc)))
        cost = np.ascontiguousarray(cost)
        out = np.zeros(nr, np.int32)
        assert lib.bt_host_lsap(cost.ctypes.data, nr, nc, out.ctypes.data) == 0
        r, c = linear_sum_assignment(cost)
        assert np.array_equal(out, c), (trial, cost)


def test_linear_assignment_equals_the_restated_one(lib):
    rng = np.random.default_rng(1)
    for trial in range(400):
        na, nb = int(rng.integers(0, 16)), int(rng.integers(0, 16))
        cost = rng.random((na, nb))
        if trial % 3 == 0 and na and nb:
            cost[rng.random((na, nb)) < 0.7] = 1.0                            # mostly above the threshold (ties after clamping)
        thresh = [0.8, 0.5, 0.7][trial % 3]
        cost = np.ascontiguousarray(cost)
        ma, mb = np.zeros(16, np.int32), np.zeros(16, np.int32)
        nua, nub = C.c_int32(), C.c_int32()
        nm = lib.bt_host_assign(cost.ctypes.data if cost.size else None, na, nb, thresh, ma.ctypes.data, mb.ctypes.data,
                                C.byref(nua), C.byref(nub))
        m, ua, ub = linear_assignment(cost, thresh)
        assert nm == len(m) and nua.value == len(ua) and nub.value == len(ub)
        if nm:
            assert np.array_equal(np.stack([ma[:nm], mb[:nm]], 1), np.asarray(m))


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("kw", [MAIN, INIT])
def test_core_equals_restated_bytetrack(lib, seed, kw):
    ref, mine = RefByteTrack(**kw), HostCore(lib, **kw)
    n_ids = 0
    for xyxy, conf in synthetic_detections(seed, n_frames=90):
        keep, ids = ref.update_with_detections(xyxy.copy(), conf.copy())
        row, tid = mine.update(xyxy, conf)
        assert np.array_equal(row, keep) and np.array_equal(tid, ids)
        n_ids = max(n_ids, ids.max() if len(ids) else 0)
    assert n_ids >= 8


def test_core_dense_scene_and_gaps(lib):
    """40 objects, long occlusions (tracks get lost, re-found, time out: max_time_lost is short here), empty frames."""
    kw = dict(track_activation_threshold=0.25, lost_track_buffer=5, minimum_matching_threshold=0.8, frame_rate=30,
              minimum_consecutive_frames=2)
    rng = np.random.default_rng(7)
    ref, mine = RefByteTrack(**kw), HostCore(lib, **kw)
    frames = synthetic_detections(11, n_frames=120, n_obj=40)
    for f, (xyxy, conf) in enumerate(frames):
        if f % 17 in (5, 6, 7, 8, 9, 10, 11):                     # blackout of a third of the objects for 7 frames
            m = rng.random(len(conf)) > 0.35
            xyxy, conf = xyxy[m], conf[m]
        if f in (50, 51):
            xyxy, conf = xyxy[:0], conf[:0]
        keep, ids = ref.update_with_detections(xyxy.copy(), conf.copy())
        row, tid = mine.update(xyxy, conf)
        assert np.array_equal(row, keep) and np.array_equal(tid, ids), f


def test_core_class_and_confidence_mask(lib):
    """The mask of main.py:189-193 applied inside the tracker == filtering before the reference tracker."""
    ref, mine = RefByteTrack(**MAIN), HostCore(lib, **MAIN)
    rng = np.random.default_rng(3)
    for xyxy, conf in synthetic_detections(5, n_frames=40, n_obj=14):
        cls = rng.integers(0, 3, len(conf)).astype(np.int32)                 # classes 0, 1 allowed; 2 masked out
        m = (conf > 0.4) & (cls < 2)
        idx = np.nonzero(m)[0]
        keep, ids = ref.update_with_detections(xyxy[m].copy(), conf[m].copy())
        row, tid = mine.update(xyxy, conf, cls, min_conf=0.4, class_mask=0b11)
        assert np.array_equal(row, idx[keep]) and np.array_equal(tid, ids)

"""K2b parity: cross-slice merge NMS vs the restated supervision box_non_max_suppression (float64,
keep mask in input order), tile gather vs numpy, and the whole slicer vs the restated InferenceSlicer."""
import numpy as np
import pytest
import torch

from hvb import _ffi
from hvb.synth import random_boxes
from oracle import supervision_restated as svr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 33, 300, 512, 513, 2000, 4096])
@pytest.mark.parametrize("agnostic", [False, True])
def test_merge_nms_mask_identical(ctx, n, agnostic):
    rng = np.random.default_rng(n + agnostic)
    boxes, scores = random_boxes(rng, n, 800, 600, 20, 140, dtype=np.float64)
    boxes += rng.integers(0, 3000, (n, 1)) * 0.0          # keep float64 path
    cls = rng.integers(0, 2, n)
    for thr in (0.1, 0.5):
        ref = svr.with_nms(boxes, scores, cls, thr, class_agnostic=agnostic)
        got = ctx.merge_nms_host(boxes, scores, cls, thr, agnostic)
        assert got.dtype == bool and np.array_equal(got, ref), (n, thr)


def test_merge_nms_degenerate_boxes(ctx):
    boxes = np.array([[10, 10, 10, 10], [10, 10, 10, 10], [0, 0, 5, 5], [0, 0, 5, 5], [1, 1, 4, 4]], np.float64)
    scores = np.array([0.9, 0.8, 0.7, 0.6, 0.5], np.float32)
    cls = np.zeros(5, np.int64)
    assert np.array_equal(ctx.merge_nms_host(boxes, scores, cls, 0.5, False), svr.with_nms(boxes, scores, cls, 0.5))


def test_merge_segments_and_gather(ctx):
    rng = np.random.default_rng(4)
    T, F, max_det = 6, 5, 16
    S = T * F
    cnt = rng.integers(0, max_det + 1, S).astype(np.int32)
    cnt[3] = 0
    xyxy = rng.uniform(0, 600, (S, max_det, 4)).astype(np.float32)
    xyxy[..., 2:] = xyxy[..., :2] + rng.uniform(5, 80, (S, max_det, 2)).astype(np.float32)
    conf = rng.permutation(np.linspace(0.3, 0.99, S * max_det)).astype(np.float32).reshape(S, max_det)
    cls = rng.integers(0, 2, (S, max_det)).astype(np.int32)
    offs = rng.integers(0, 2000, (S, 2)).astype(np.float32)
    dev = lambda a: torch.from_numpy(a).cuda()
    g_xyxy, g_conf, g_cls, g_slot, seg = ctx.gather_tiles(dev(xyxy), dev(conf), dev(cls), dev(cnt), dev(offs), T, max_det)
    seg = seg.cpu().numpy()
    total = int(cnt.sum())
    assert seg[-1] == total
    ref_x, ref_c, ref_k, ref_seg = [], [], [], [0]
    for s in range(S):
        k = cnt[s]
        ref_x.append(svr.move_boxes(xyxy[s, :k], offs[s].astype(np.int64)))
        ref_c.append(conf[s, :k]); ref_k.append(cls[s, :k])
        if s % T == T - 1:
            ref_seg.append(ref_seg[-1] + int(cnt[s - T + 1:s + 1].sum()))
    assert np.array_equal(seg, np.array(ref_seg))
    ref_x = np.concatenate(ref_x)
    assert ref_x.dtype == np.float64
    assert np.array_equal(g_xyxy[:total].cpu().numpy(), ref_x)
    assert np.array_equal(g_conf[:total].cpu().numpy(), np.concatenate(ref_c))
    keep = ctx.merge_nms(g_xyxy, g_conf, g_cls, torch.from_numpy(seg).cuda(), F, total, 0.3, False).cpu().numpy().astype(bool)
    for f in range(F):
        lo, hi = seg[f], seg[f + 1]
        ref = svr.with_nms(ref_x[lo:hi], np.concatenate(ref_c)[lo:hi], np.concatenate(ref_k)[lo:hi], 0.3)
        assert np.array_equal(keep[lo:hi], ref)


def test_detections_with_nms_surface(ctx):
    from hvb.detections import Detections
    rng = np.random.default_rng(2)
    boxes, scores = random_boxes(rng, 200, 500, 500, 30, 120)
    cls = rng.integers(0, 2, 200)
    d = Detections(xyxy=boxes, confidence=scores, class_id=cls)
    out = d.with_nms(0.4)
    ref = svr.with_nms(boxes, scores, cls, 0.4)
    assert np.array_equal(out.xyxy, boxes[ref]) and np.array_equal(out.class_id, cls[ref])

"""CPU: independent cross-checks of the RESTATED (parity-unpinned) oracle pieces against real libraries that are
installed here.  ultralytics and supervision themselves are absent, so these do not pin the
restatements to their originals; they do pin the arithmetic they share with well-known library routines:

  * supervision `box_iou_batch` / `box_non_max_suppression`  vs  torchvision.ops.box_iou / nms / batched_nms
  * ultralytics LetterBox resize + border                      vs  cv2.resize + cv2.copyMakeBorder driven independently
  * ultralytics `scale_boxes` / letterbox geometry              vs  an exact rational-arithmetic inverse of the letterbox
  * ultralytics DFL decode (softmax expectation, dist2bbox)     vs  a float64 numpy evaluation of the same formula
  * ByteTrack `linear_assignment`                               vs  brute-force optimum on small problems
"""
import itertools
from fractions import Fraction

import cv2
import numpy as np
import pytest
import torch
import torchvision

from oracle import bytetrack_restated as bt
from oracle import supervision_restated as svr
from oracle import ultralytics_restated as ur


def _boxes(rng, n, w=1920, h=1080, smin=10, smax=200):
    cx, cy = rng.uniform(0, w, n), rng.uniform(0, h, n)
    bw, bh = rng.uniform(smin, smax, n), rng.uniform(smin, smax, n)
    return np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)


@pytest.mark.parametrize("n,m", [(1, 1), (7, 13), (64, 50), (300, 300)])
def test_box_iou_batch_matches_torchvision(n, m):
    rng = np.random.default_rng(n * 1000 + m)
    a, b = _boxes(rng, n), _boxes(rng, m)
    got = svr.box_iou_batch(a, b)
    ref = torchvision.ops.box_iou(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-12


@pytest.mark.parametrize("n", [2, 33, 300, 1000])
@pytest.mark.parametrize("thr", [0.1, 0.5])
def test_restated_supervision_nms_keeps_what_torchvision_keeps(n, thr):
    """Class-agnostic and class-aware: same greedy suppression as torchvision (distinct scores, fp64 boxes)."""
    rng = np.random.default_rng(n)
    boxes = _boxes(rng, n, 800, 600, 20, 120)
    scores = rng.permutation(np.linspace(0.2, 0.99, n))
    cls = rng.integers(0, 3, n)
    keep = svr.with_nms(boxes, scores.astype(np.float32), None, thr, class_agnostic=True)
    ref = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
    assert np.array_equal(np.nonzero(keep)[0], np.sort(ref))
    keep = svr.with_nms(boxes, scores.astype(np.float32), cls, thr, class_agnostic=False)
    ref = torchvision.ops.batched_nms(torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(cls), thr).numpy()
    assert np.array_equal(np.nonzero(keep)[0], np.sort(ref))


@pytest.mark.parametrize("h,w,imgsz", [(1080, 1920, 1280), (720, 1280, 640), (640, 624, 640), (137, 640, 640), (2160, 3840, 1280)])
def test_letterbox_is_resize_plus_centered_border(h, w, imgsz):
    """LetterBox(auto=True): scale-to-fit, INTER_LINEAR resize, pad to a multiple of 32 split evenly, value 114."""
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    out = ur.letterbox(img, imgsz, auto=True)
    r = min(imgsz / h, imgsz / w)
    nw, nh = int(round(w * r)), int(round(h * r))
    dw, dh = ((imgsz - nw) % 32) / 2, ((imgsz - nh) % 32) / 2
    res = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR) if (nw, nh) != (w, h) else img
    ref = cv2.copyMakeBorder(res, int(round(dh - 0.1)), int(round(dh + 0.1)), int(round(dw - 0.1)), int(round(dw + 0.1)),
                             cv2.BORDER_CONSTANT, value=(114, 114, 114))
    assert out.shape == ref.shape and out.shape[0] % 32 == 0 and out.shape[1] % 32 == 0
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("img0", [(1080, 1920), (720, 1280), (2160, 3840), (601, 333)])
def test_scale_boxes_inverts_the_letterbox(img0):
    """Boxes mapped into letterbox coordinates with exact rational arithmetic come back to within 1e-3 px."""
    h, w = img0
    g = ur.letterbox(np.zeros((h, w, 3), np.uint8), 1280, auto=True).shape[:2]
    gain, px, py = ur.scale_boxes_geometry(g, img0)
    rng = np.random.default_rng(h)
    b = _boxes(rng, 50, w, h, 5, 300)
    b[:, [0, 2]] = b[:, [0, 2]].clip(0, w); b[:, [1, 3]] = b[:, [1, 3]].clip(0, h)
    fwd = np.array([[float(Fraction(v) * Fraction(gain) + (px if k % 2 == 0 else py)) for k, v in enumerate(row)] for row in b])
    back = ur.scale_boxes(g, torch.from_numpy(fwd.astype(np.float32)), img0).numpy()
    assert np.abs(back - b).max() <= 2e-3


def test_dfl_decode_matches_float64_formula():
    """Detect._inference: softmax expectation over 16 bins per side, dist2bbox (xywh), x stride, sigmoid classes."""
    rng = np.random.default_rng(0)
    hw = (64, 96)
    lv = [(hw[0] // s, hw[1] // s) for s in (8, 16, 32)]
    levels = [torch.from_numpy(rng.normal(0, 2, (1, 64 + 3, a, b)).astype(np.float32)) for a, b in lv]
    got = ur.decode_head(levels, 3).numpy()[0]                       # [4+nc, A]
    cols = []
    for (a, b), s, t in zip(lv, (8, 16, 32), levels):
        x = t[0].numpy().astype(np.float64).reshape(67, -1)
        d = x[:64].reshape(4, 16, -1)
        p = np.exp(d - d.max(1, keepdims=True)); p /= p.sum(1, keepdims=True)
        dist = (p * np.arange(16)[None, :, None]).sum(1)             # [4, a*b]  l, t, r, b
        gy, gx = np.divmod(np.arange(a * b), b)
        ax, ay = gx + 0.5, gy + 0.5
        x1, y1, x2, y2 = ax - dist[0], ay - dist[1], ax + dist[2], ay + dist[3]
        box = np.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1]) * s
        cols.append(np.vstack([box, 1 / (1 + np.exp(-x[64:]))]))
    ref = np.hstack(cols)
    assert got.shape == ref.shape
    assert np.abs(got[:4] - ref[:4]).max() <= 1e-3 and np.abs(got[4:] - ref[4:]).max() <= 1e-6


@pytest.mark.parametrize("n,m", [(1, 1), (3, 5), (6, 4), (7, 7)])
def test_linear_assignment_is_optimal_on_small_problems(n, m):
    rng = np.random.default_rng(10 * n + m)
    for _ in range(20):
        cost = rng.uniform(0, 1, (n, m))
        thr = 0.6
        matches, ua, ub = bt.linear_assignment(cost, thr)
        assert len(matches) + len(ua) == n and len(matches) + len(ub) == m
        assert all(cost[i, j] <= thr for i, j in matches)
        c = np.where(cost > thr, thr + 1e-4, cost)                   # what the restated function minimises
        k = min(n, m)
        best = min(sum(c[i, j] for i, j in zip(rows, cols))
                   for rows in itertools.combinations(range(n), k) for cols in itertools.permutations(range(m), k))
        # the matched pairs are exactly the under-threshold part of an optimal assignment: their cost plus the
        # clamped cost of one pair per remaining min(n,m) slot equals the brute-force optimum
        got = sum(c[i, j] for i, j in matches) + (k - len(matches)) * (thr + 1e-4)
        assert abs(got - best) <= 1e-12

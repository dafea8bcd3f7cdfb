"""K3a parity: bit-exact HSV/LAB vs cv2.cvtColor over the full 2^24 colour cube; exact integer
histograms / sums / counts and the 49-d feature vs the reference arithmetic (oracle.team_reference,
itself pinned against the real reference); crop geometry vs sv.crop_image."""
import hashlib
import os
import sys

import cv2
import numpy as np
import pytest
import torch

from hvb import _ffi
from hvb.synth import pack_crops, random_frames
from oracle import team_reference as tr
from oracle.supervision_restated import crop_image

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))


def descs(crops):
    buf, d = pack_crops(crops)
    cd = np.zeros((len(crops),), _ffi.CROP_DESC)
    cd["offset"], cd["pitch"], cd["h"], cd["w"] = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
    return buf, cd


def test_full_colour_cube_bit_exact(ctx):
    b, g, r = np.meshgrid(*[np.arange(256, dtype=np.uint8)] * 3, indexing="ij")
    cube = np.ascontiguousarray(np.stack([b, g, r], -1).reshape(4096, 4096, 3))
    hsv, lab = ctx.cvt_hsv_lab(torch.from_numpy(cube).cuda())
    hsv, lab = hsv.cpu().numpy(), lab.cpu().numpy()
    assert np.array_equal(hsv, cv2.cvtColor(cube, cv2.COLOR_BGR2HSV))
    assert np.array_equal(lab, cv2.cvtColor(cube, cv2.COLOR_BGR2LAB))
    assert hashlib.sha256(hsv.tobytes()).hexdigest() == "cc4c8f3a2064ffaed3776170c4dfa02c90011b02dbc7ece07b7fced54069ad55"
    assert hashlib.sha256(lab.tobytes()).hexdigest() == "6777b2103b2347e79cfcaf30f90002ede0142304b76abfd6100310e5ea68480c"


def check_against_oracle(ctx, crops, roi_mode=_ffi.ROI_HYBRID):
    buf, cd = descs(crops)
    feat, raw = ctx.color_features_host(buf, cd, roi_mode, want_raw=True)
    for i, c in enumerate(crops):
        roi = tr.jersey_region(c) if roi_mode == _ffi.ROI_HYBRID else c
        ref = tr.color_stats_raw(roi)
        assert raw["n"][i] == ref["n"]
        assert np.array_equal(raw["hist"][i], ref["hist"]), i
        assert np.array_equal(raw["counts"][i], ref["counts"]), i
        assert np.array_equal(raw["sums"][i], ref["sums"]), i
        assert np.array_equal(raw["sumsq"][i], ref["sumsq"]), i
    if roi_mode == _ffi.ROI_HYBRID:
        ref_f = tr.color_features(crops)
        # histogram / ratio features are exact float32->float64 / integer ratios; mean/std within 1e-9 rel
        assert np.array_equal(feat[:, :34], ref_f[:, :34])
        assert np.array_equal(feat[:, 46:], ref_f[:, 46:])
        np.testing.assert_allclose(feat[:, 34:46], ref_f[:, 34:46], rtol=1e-9, atol=1e-12)
    return feat, raw


def test_random_rois_max_contention(ctx):
    rng = np.random.default_rng(0)
    crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in
             [(1, 1), (2, 3), (39, 50), (40, 19), (40, 20), (100, 40), (150, 70), (250, 110), (333, 97), (600, 400)]]
    check_against_oracle(ctx, crops)
    check_against_oracle(ctx, crops, _ffi.ROI_WHOLE)


def test_flat_colour_rois_single_bin(ctx):
    crops = [np.full((120, 60, 3), v, np.uint8) for v in (0, 17, 128, 255)]
    crops += [np.broadcast_to(np.array(c, np.uint8), (90, 45, 3)).copy() for c in ((235, 235, 235), (40, 40, 200), (0, 255, 0))]
    feat, raw = check_against_oracle(ctx, crops)
    assert (feat[:4, 37:40] == 0).all() and (feat[:4, 43:46] == 0).all()      # std of a constant ROI is exactly 0


def test_noncontiguous_views_and_frame_resident_path(ctx):
    frames = random_frames(1, 2, 540, 960)
    rng = np.random.default_rng(2)
    boxes = np.stack([rng.uniform(-20, 900, 24), rng.uniform(-20, 500, 24), rng.uniform(30, 1000, 24), rng.uniform(30, 560, 24)], 1).astype(np.float32)
    boxes[:, 2] = np.maximum(boxes[:, 2], boxes[:, 0] + 1)
    boxes[:, 3] = np.maximum(boxes[:, 3], boxes[:, 1] + 1)
    boxes[3] = [10.5, 20.5, 61.5, 120.5]           # half-to-even rounding
    boxes[4] = [100.2, 50.7, 100.4, 90.0]          # zero-width after rounding
    fidx = (np.arange(24) % 2).astype(np.int32)
    crops = [crop_image(frames[f], b) for f, b in zip(fidx, boxes)]
    assert not crops[0].flags["C_CONTIGUOUS"]
    cd_dev = ctx.crops_from_boxes(torch.from_numpy(boxes).cuda(), torch.from_numpy(fidx).cuda(), 540, 960)
    cd = cd_dev.cpu().numpy().view(_ffi.CROP_DESC)[:24]
    for i, c in enumerate(crops):
        assert (cd["h"][i], cd["w"][i]) == c.shape[:2], (i, boxes[i])
    keep = [i for i, c in enumerate(crops) if c.size > 0 and tr.jersey_region(c).size > 0]
    f_dev, raw_dev = ctx.color_features(torch.from_numpy(frames).cuda(), cd_dev, 24, want_raw=True)
    f_dev = f_dev.cpu().numpy()
    ref = tr.color_features([crops[i] for i in keep])
    assert np.array_equal(f_dev[keep][:, :34], ref[:, :34])
    np.testing.assert_allclose(f_dev[keep], ref, rtol=1e-9, atol=1e-12)
    empty = [i for i in range(24) if i not in keep]
    assert len(empty) >= 1 and np.isnan(f_dev[empty]).all()
    # packed-crop path gives the same bits as the frame-resident path
    f_host = ctx.color_features_host(*descs([crops[i] for i in keep]))
    assert np.array_equal(f_host, f_dev[keep])


def test_golden_reference_colour_features(ctx):
    from make_golden import golden_crops
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "team_reference.npz"))
    _, crops, *_ = golden_crops()
    feat = ctx.color_features_host(*descs(crops))
    assert np.array_equal(feat[:, :34], gold["color"][:, :34])
    assert np.array_equal(feat[:, 46:], gold["color"][:, 46:])
    np.testing.assert_allclose(feat, gold["color"], rtol=1e-9, atol=1e-12)


def test_simple_rule_roi(ctx):
    from make_golden import golden_crops
    from hvb.team import TeamClassifier
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "team_reference.npz"))
    _, crops, *_ = golden_crops()
    tc = TeamClassifier(device="cuda:0", use_hybrid=False)
    out = [tc.classify_jersey(c) for c in crops]
    assert np.array_equal(np.array([t for t, _ in out]), gold["simple_team"])
    np.testing.assert_allclose(np.array([c for _, c in out]), gold["simple_conf"], rtol=0, atol=1e-12)


def test_large_property_histogram_mass(ctx):
    """Size-independent property at full 4K size: every histogram sums to the ROI pixel count."""
    frames = random_frames(5, 1, 2160, 3840)
    cd = np.zeros((1,), _ffi.CROP_DESC)
    cd["offset"], cd["pitch"], cd["h"], cd["w"] = 0, 3840 * 3, 2160, 3840
    _, raw = ctx.color_features_host(frames.reshape(-1), cd, _ffi.ROI_WHOLE, want_raw=True)
    n = 2160 * 3840
    assert raw["n"][0] == n
    assert raw["hist"][0][:18].sum() == n and raw["hist"][0][18:26].sum() == n and raw["hist"][0][26:].sum() == n
    hsv = cv2.cvtColor(frames[0], cv2.COLOR_BGR2HSV).reshape(-1, 3).astype(np.uint64)
    assert np.array_equal(raw["sums"][0][:3], hsv.sum(0))

"""K3b parity: Pillow-exact resize + ToTensor + Normalize, bit-exact vs the reference's own
transforms.Compose (team_hybrid.py:31-36) on jersey ROIs; deep features vs the per-crop CPU forward."""
import os
import sys

import numpy as np
import pytest
import torch

from hvb import _ffi
from hvb.synth import pack_crops
from oracle import cv_exact as cx
from oracle import team_reference as tr

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))


def descs(crops):
    buf, d = pack_crops(crops)
    cd = np.zeros((len(crops),), _ffi.CROP_DESC)
    cd["offset"], cd["pitch"], cd["h"], cd["w"] = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
    return buf, cd


SHAPES = [(1, 1), (5, 3), (39, 50), (40, 20), (75, 33), (125, 66), (250, 110), (128, 64), (200, 64), (128, 100),
          (300, 2), (201, 2), (200, 2), (401, 4), (64, 128), (17, 200), (500, 300), (101, 1), (1000, 640), (1500, 900)]


def test_preprocess_bit_exact_whole_roi(ctx):
    rng = np.random.default_rng(0)
    crops = []
    for (h, w) in SHAPES:
        crops.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        crops.append((rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8))
    buf, cd = descs(crops)
    out, valid, u8 = ctx.mnv3_preprocess(torch.from_numpy(buf).cuda(), ctx.struct_to_device(cd), len(crops), _ffi.ROI_WHOLE, want_u8=True)
    out, valid, u8 = out.cpu().numpy(), valid.cpu().numpy(), u8.cpu().numpy()
    pp = tr.make_preprocess()
    for i, c in enumerate(crops):
        assert valid[i] == 1
        assert np.array_equal(u8[i], cx.pil_resize_bilinear(c, 64, 128)), c.shape
        assert np.array_equal(out[i], pp(c).numpy()), c.shape


def test_preprocess_jersey_roi_and_golden(ctx):
    from make_golden import golden_crops
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "team_reference.npz"))
    _, crops, *_ = golden_crops()
    out, valid = ctx.mnv3_preprocess_host(*descs(crops))
    assert (valid == 1).all()
    assert np.array_equal(out, gold["preprocessed"])


def test_empty_roi_yields_zero_row_flag(ctx):
    crops = [np.zeros((0, 10, 3), np.uint8), np.full((50, 30, 3), 9, np.uint8), np.zeros((10, 0, 3), np.uint8)]
    out, valid = ctx.mnv3_preprocess_host(*descs(crops))
    assert list(valid) == [0, 1, 0] and (out[0] == 0).all() and (out[2] == 0).all()


def test_deep_features_match_reference_forward(ctx):
    from make_golden import golden_crops
    from hvb.hybrid import HybridTeamClassifier
    from hvb.models import build_trunk
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "team_reference.npz"))
    _, crops, *_ = golden_crops()
    clf = HybridTeamClassifier(device="cuda:0", trunk=build_trunk(0))
    deep = clf.extract_deep_features(crops)
    g = gold["deep"]
    scale = np.abs(g).max(axis=1, keepdims=True)
    assert (np.abs(deep - g) <= 1e-3 * scale).all(), float((np.abs(deep - g) / scale).max())
    # well-conditioned (BN-calibrated) trunk shared by both paths: O(1) features, same tolerance
    trunk = build_trunk(0, calibrate=True)
    ref = tr.deep_features(trunk, crops[:16])
    clf2 = HybridTeamClassifier(device="cuda:0", trunk=trunk)
    got = clf2.extract_deep_features(crops[:16])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert scale.min() > 0.1
    assert (np.abs(got - ref) <= 1e-3 * scale).all()

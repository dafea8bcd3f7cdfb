"""CPU: oracle/segmentation_reference.py (the GrabCut-free SegmentationTeamClassifier path)
against the golden outputs of the REAL reference class (tests/golden/segmentation_reference.npz) and, in the build
container, against a live import of /root/reference/hockey/common/team_segmentation.py."""
import os
import sys

import numpy as np
import pytest

from oracle import reference_loader as rl
from oracle import segmentation_reference as sr

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import golden_crops  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "segmentation_reference.npz")


@pytest.fixture(scope="module")
def data():
    _, crops, labels, _, tids = golden_crops()
    return crops, tids, int((labels >= 0).sum()), np.load(GOLD)


def _run_oracle(crops, tids, n_fit):
    out = {}
    masks = [sr.fallback_mask(*c.shape[:2]) for c in crops]
    out["mask_rect"] = np.array([[np.argmax(m.any(1)) if m.any() else 0, m.any(1).sum(), np.argmax(m.any(0)) if m.any() else 0,
                                  m.any(0).sum(), m.sum()] for m in masks])
    out["features"] = np.array([sr.feature_row(sr.extract_jersey_colors(c, m)) for c, m in zip(crops, masks)], np.float64)
    cj = [sr.classify_single_jersey(c) for c in crops]
    out["single_team"] = np.array([t for t, _ in cj])
    out["single_conf"] = np.array([c for _, c in cj], np.float64)
    out["predict_unfitted"] = sr.SegmentationReference().predict(list(crops[:n_fit]), tids[:n_fit])
    clf = sr.SegmentationReference()
    clf.fit(list(crops[:n_fit]))
    out["centers"] = clf.kmeans.cluster_centers_
    per = n_fit // 4
    out["predict"] = np.concatenate([clf.predict(list(crops[f * per:(f + 1) * per]), tids[f * per:(f + 1) * per]) for f in range(4)])
    out["predict_no_ids"] = clf.predict(list(crops[:n_fit]))
    return out


def test_restatement_matches_the_golden_outputs_of_the_real_reference(data):
    crops, tids, n_fit, gold = data
    got = _run_oracle(crops, tids, n_fit)
    assert set(got) == set(gold.files)
    for k in gold.files:
        if k == "centers":
            assert np.allclose(got[k], gold[k], rtol=0, atol=1e-9), k
        else:
            assert np.array_equal(got[k], gold[k]), k                     # bit-exact, float64 values included
    assert (gold["features"][:, 0] > 0).any() and set(gold["predict"]) == {0, 1}


def test_the_uint8_wrap_in_the_white_test_is_reproduced():
    """a = 120 is 8 below neutral: |a-128| < 10 mathematically, but the reference's uint8 subtraction wraps."""
    import cv2
    for bgr in ((235, 235, 235), (255, 255, 255), (250, 245, 235), (235, 245, 250)):
        crop = np.full((60, 40, 3), bgr, np.uint8)
        lab = cv2.cvtColor(crop[:1, :1], cv2.COLOR_BGR2LAB)[0, 0]
        f = sr.extract_jersey_colors(crop, sr.fallback_mask(60, 40))
        want = float(lab[0] > 200 and 128 <= lab[1] < 138 and 128 <= lab[2] < 138)
        assert f["is_white"] == want, (bgr, lab)


@pytest.mark.skipif(not rl.available(), reason="reference tree not mounted (GPU box)")
def test_restatement_matches_a_live_import_of_the_reference(data):
    from make_golden_segmentation import load_real, run_real
    crops, tids, n_fit, gold = data
    live = run_real(load_real(), crops, tids, n_fit)
    got = _run_oracle(crops, tids, n_fit)
    for k in gold.files:
        assert np.allclose(live[k], gold[k], rtol=0, atol=1e-9), k
        assert np.allclose(got[k], live[k], rtol=0, atol=1e-9), k
    rng = np.random.default_rng(3)                                        # extra random crops, features only
    seg = load_real()
    clf = seg.SegmentationTeamClassifier()
    for _ in range(30):
        h, w = int(rng.integers(20, 260)), int(rng.integers(10, 120))
        base = rng.integers(0, 256, 3)
        crop = np.clip(base + rng.normal(0, 20, (h, w, 3)), 0, 255).astype(np.uint8)
        a = clf.extract_jersey_colors(crop, clf.segment_player(crop))
        b = sr.extract_jersey_colors(crop, sr.fallback_mask(h, w))
        assert sr.feature_row(a) == sr.feature_row(b)


def test_host_side_feature_arithmetic_of_the_product_matches_the_golden_features(data):
    """hvb.team_segmentation.features_from_raw (host logic of the product, no GPU): exact integer statistics, here
    produced by real cv2 conversions instead of the K3c kernel, must give the golden float64 features bit for bit."""
    import cv2
    from hvb import _ffi
    from hvb.team_segmentation import DEFAULTS, features_from_raw
    crops, _, _, gold = data
    raw = np.zeros(len(crops), _ffi.JERSEY_RAW)
    for i, c in enumerate(crops):
        px = c[sr.fallback_mask(*c.shape[:2])]
        raw[i]["n"] = len(px)
        if len(px) == 0:
            continue
        hsv = cv2.cvtColor(px.reshape(-1, 1, 3), cv2.COLOR_BGR2HSV).reshape(-1, 3).astype(np.int64)
        lab = cv2.cvtColor(px.reshape(-1, 1, 3), cv2.COLOR_BGR2LAB).reshape(-1, 3).astype(np.int64)
        white = (lab[:, 0] > 200) & (lab[:, 1] >= 128) & (lab[:, 1] < 138) & (lab[:, 2] >= 128) & (lab[:, 2] < 138)
        raw[i]["white"] = white.sum()
        raw[i]["hue_hist"] = np.bincount(hsv[~white, 0] // 10, minlength=18)
        raw[i]["sat_colored"], raw[i]["sat_all"], raw[i]["val_all"] = hsv[~white, 1].sum(), hsv[:, 1].sum(), hsv[:, 2].sum()
    got = features_from_raw(raw)
    assert np.array_equal(got, gold["features"])
    # fewer than 100 masked pixels -> the reference's defaults; 100+ pixels but <= 50 non-white -> hue 0, mean S of ALL pixels
    few = np.zeros(2, _ffi.JERSEY_RAW)
    few[0]["n"], few[0]["white"], few[0]["sat_all"], few[0]["val_all"] = 99, 99, 990, 9900
    few[1]["n"], few[1]["white"], few[1]["sat_all"], few[1]["sat_colored"], few[1]["val_all"] = 120, 70, 1200, 1100, 24000
    few[1]["hue_hist"][5] = 50
    out = features_from_raw(few)
    assert tuple(out[0]) == DEFAULTS
    assert tuple(out[1]) == (70 / 120, 0.0, 1200 / 120, 200.0)

"""SegmentationTeamClassifier colour features without GrabCut, on the GPU
(hvb_jersey_color_stats) vs the oracle (oracle/segmentation_reference.py, pinned to the real reference by
tests/golden/segmentation_reference.npz).  Counts and sums are integers, so every feature must be bit-identical."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import golden_crops  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "segmentation_reference.npz")


@pytest.fixture(scope="module")
def data():
    frames, crops, labels, _, tids = golden_crops()
    return frames, crops, tids, int((labels >= 0).sum()), np.load(GOLD)


def _raw_oracle(crop, mask):
    """hvb_jersey_raw fields from real cv2 conversions."""
    import cv2
    px = crop[mask]
    if len(px) == 0:
        return 0, 0, np.zeros(18, np.int64), 0, 0, 0
    hsv = cv2.cvtColor(px.reshape(-1, 1, 3), cv2.COLOR_BGR2HSV).reshape(-1, 3).astype(np.int64)
    lab = cv2.cvtColor(px.reshape(-1, 1, 3), cv2.COLOR_BGR2LAB).reshape(-1, 3).astype(np.int64)
    white = (lab[:, 0] > 200) & (lab[:, 1] >= 128) & (lab[:, 1] < 138) & (lab[:, 2] >= 128) & (lab[:, 2] < 138)
    hist = np.bincount(hsv[~white, 0] // 10, minlength=18)
    return len(px), int(white.sum()), hist, int(hsv[~white, 1].sum()), int(hsv[:, 1].sum()), int(hsv[:, 2].sum())


def test_raw_statistics_are_bit_exact(ctx, data):
    from hvb import _ffi
    from hvb.synth import pack_crops
    from oracle import segmentation_reference as sr
    _, crops, _, _, _ = data
    rng = np.random.default_rng(11)
    extra = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((200, 90), (255, 255), (37, 23), (5, 3), (600, 300))]
    extra += [np.full((120, 60, 3), c, np.uint8) for c in ((255, 255, 255), (250, 245, 235), (0, 0, 0), (40, 40, 200))]
    allc = list(crops) + extra
    buf, desc = pack_crops(allc)
    cd = np.zeros((len(allc),), _ffi.CROP_DESC)
    cd["offset"], cd["pitch"], cd["h"], cd["w"] = desc[:, 0], desc[:, 1], desc[:, 2], desc[:, 3]
    before = ctx.launch_count()
    for mode in (_ffi.ROI_SEGMENT, _ffi.ROI_WHOLE, _ffi.ROI_HYBRID):
        raw = ctx.jersey_color_stats_host(buf, cd, mode)
        for i, c in enumerate(allc):
            h, w = c.shape[:2]
            if mode == _ffi.ROI_SEGMENT:
                mask = sr.fallback_mask(h, w)
            elif mode == _ffi.ROI_WHOLE:
                mask = np.ones((h, w), bool)
            else:
                mask = np.zeros((h, w), bool)
                if h < 40 or w < 20:
                    mask[:] = True
                else:
                    mask[int(h * 0.1):int(h * 0.6), int(w * 0.2):int(w * 0.8)] = True
            n, white, hist, sc, sa, va = _raw_oracle(c, mask)
            r = raw[i]
            assert (int(r["n"]), int(r["white"]), int(r["sat_colored"]), int(r["sat_all"]), int(r["val_all"])) == (n, white, sc, sa, va), (mode, i)
            assert np.array_equal(r["hue_hist"].astype(np.int64), hist), (mode, i)
    assert ctx.launch_count() >= before + 3


def test_features_and_rule_match_the_golden_reference_outputs(ctx, data):
    from hvb import SegmentationTeamClassifier
    from oracle import segmentation_reference as sr
    _, crops, _, _, gold = data
    clf = SegmentationTeamClassifier("cuda:0")
    feats, npx = clf.jersey_features(crops)
    assert np.array_equal(feats, gold["features"])                        # float64, bit for bit
    assert np.array_equal(npx, gold["mask_rect"][:, 4])
    for i in (0, 1, 40, 41, 46):
        f = clf.extract_jersey_colors(crops[i], clf.segment_player(crops[i]))
        assert [f["is_white"], f["dominant_hue"], f["saturation"], f["brightness"]] == list(gold["features"][i])
        assert np.array_equal(clf.segment_player(crops[i]), sr.fallback_mask(*crops[i].shape[:2]))
        t, c = clf.classify_single_jersey(crops[i])
        assert (t, c) == (gold["single_team"][i], gold["single_conf"][i])
    with pytest.raises(NotImplementedError):
        m = clf.segment_player(crops[0]); m[m.shape[0] // 2, 0] = True
        clf.extract_jersey_colors(crops[0], m)


def test_fit_predict_and_vote_match_the_reference(ctx, data, capsys):
    from hvb import SegmentationTeamClassifier
    _, crops, tids, n_fit, gold = data
    clf = SegmentationTeamClassifier("cuda:0")
    assert np.array_equal(clf.predict(list(crops[:n_fit]), tids[:n_fit]), gold["predict_unfitted"])
    clf = SegmentationTeamClassifier("cuda:0", visualize_segmentation=True)
    clf.fit(list(crops[:n_fit]))
    assert "avg white ratio" in capsys.readouterr().out
    assert np.allclose(clf.kmeans.cluster_centers_, gold["centers"], rtol=0, atol=1e-9)
    per = n_fit // 4
    got = np.concatenate([clf.predict(list(crops[f * per:(f + 1) * per]), tids[f * per:(f + 1) * per]) for f in range(4)])
    assert np.array_equal(got, gold["predict"])
    assert set(clf.get_segmentation_masks([int(t) for t in tids[:per]])) == {int(t) for t in tids[:per]}
    clf.player_history.clear()
    assert np.array_equal(clf.predict(list(crops[:n_fit])), gold["predict_no_ids"])
    assert len(clf.predict([])) == 0


def test_device_resident_frames_give_the_same_features(ctx, data):
    """Boxes into frames already in HBM (crops_from_boxes geometry) == host crops cut by sv.crop_image."""
    from hvb import SegmentationTeamClassifier
    from hvb.synth import rink_clip
    from oracle.supervision_restated import crop_image
    frames, boxes, _, _ = rink_clip(7, 4, 1080, 1920, 10)
    xyxy = np.concatenate([b for b in boxes]).astype(np.float32)
    fidx = np.repeat(np.arange(4), [len(b) for b in boxes]).astype(np.int32)
    clf = SegmentationTeamClassifier("cuda:0")
    dev = clf.jersey_features_from_frames(torch.from_numpy(np.stack(frames)).cuda(), xyxy, fidx)
    host, _ = clf.jersey_features([crop_image(frames[f], b) for f, b in zip(fidx, xyxy)])
    assert np.array_equal(dev, host)
    clf.fit([crop_image(frames[f], b) for f, b in zip(fidx, xyxy)])
    a = clf.predict_from_frames(torch.from_numpy(np.stack(frames)).cuda(), xyxy, fidx)
    b = clf.predict([crop_image(frames[f], b) for f, b in zip(fidx, xyxy)])
    assert np.array_equal(a, b) and set(a) == {0, 1}


def test_router_opt_in_and_cascade(ctx, data, capsys):
    """TeamClassifier(use_segmentation="rectangle") routes like the reference's default flags do when
    team_segmentation imports (team.py:47-56, 141-154, 227-238), and a failure lands on the simple HSV rule."""
    from hvb import TeamClassifier
    _, crops, tids, n_fit, gold = data
    tc = TeamClassifier("cuda:0", use_segmentation="rectangle")
    assert tc.use_segmentation and not tc.use_hybrid and not hasattr(tc, "hybrid_classifier")
    tc.fit(list(crops[:n_fit]), positions=[(0.0, 0.0)] * n_fit)
    per = n_fit // 4
    got = np.concatenate([tc.predict(list(crops[f * per:(f + 1) * per]), tids[f * per:(f + 1) * per]) for f in range(4)])
    assert np.array_equal(got, gold["predict"])
    masks = tc.get_segmentation_masks([int(tids[0])])
    assert set(masks) == {int(tids[0])} and masks[int(tids[0])].dtype == bool
    assert TeamClassifier("cuda:0").get_segmentation_masks([1]) is None            # default flags: hybrid route

    tc2 = TeamClassifier("cuda:0", use_segmentation="rectangle")

    def boom(*a, **k):
        raise RuntimeError("planted failure")
    tc2.segmentation_classifier.predict = boom
    team_gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "team_reference.npz"))
    out = tc2.predict(list(crops[:n_fit]), tids[:n_fit])
    assert "Segmentation prediction failed: planted failure" in capsys.readouterr().out
    assert not tc2.use_segmentation and not tc2.use_hybrid
    assert np.array_equal(out, team_gold["simple_predict"])                          # the simple rule of team.py

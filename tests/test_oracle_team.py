"""Pin oracle/team_reference.py against the reference's own team_hybrid.py / team.py:
 - always: against tests/golden/team_reference.npz (outputs of the real reference, see make_golden.py)
 - when /root/reference is mounted (build container): against a live import as well."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import golden_crops  # noqa: E402

from oracle import reference_loader as rl  # noqa: E402
from oracle import team_reference as tr  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "team_reference.npz"))


@pytest.fixture(scope="module")
def data():
    torch.set_num_threads(1)
    frames, crops, labels, positions, tids = golden_crops()
    from hvb.models import build_trunk
    return dict(crops=crops, labels=labels, positions=positions, tids=tids, trunk=build_trunk(0))


def test_color_features_match_reference(data):
    f = tr.color_features(data["crops"])
    assert f.shape == (47, 49)
    assert np.array_equal(f, GOLD["color"])


def test_jersey_rect(data):
    shapes = np.array([tr.jersey_region(c).shape[:2] for c in data["crops"]])
    assert np.array_equal(shapes, GOLD["jersey_shapes"])


def test_preprocess_matches_reference(data):
    assert np.array_equal(tr.preprocess_rois(data["crops"]), GOLD["preprocessed"])


def test_deep_features_match_reference(data):
    d = tr.deep_features(data["trunk"], data["crops"]).astype(np.float32)
    g = GOLD["deep"]
    scale = np.abs(g).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(d - g) <= 1e-5 * scale).all()


def test_fit_affinity_and_predict_match_reference(data):
    n_fit = int((data["labels"] >= 0).sum())
    ref = tr.HybridReference(data["trunk"])
    ref.fit(data["crops"][:n_fit])
    np.testing.assert_allclose(ref.scaler.mean_, GOLD["scaler_mean"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(ref.affinity_matrix_, GOLD["affinity"], rtol=1e-6, atol=1e-300)
    per = n_fit // 4
    preds = np.concatenate([ref.predict(data["crops"][f * per:(f + 1) * per], data["tids"][f * per:(f + 1) * per]) for f in range(4)])
    assert np.array_equal(preds, GOLD["predict"])


def test_simple_rule_matches_reference(data):
    out = [tr.simple_jersey_rule(c) for c in data["crops"]]
    assert np.array_equal(np.array([t for t, _ in out]), GOLD["simple_team"])
    np.testing.assert_allclose(np.array([c for _, c in out]), GOLD["simple_conf"], rtol=0, atol=1e-12)


@pytest.mark.skipif(not rl.available(), reason="/root/reference not mounted (GPU box)")
def test_live_reference_import(data):
    hyb = rl.make_hybrid(seed=0)
    crops = data["crops"][:12]
    assert np.array_equal(hyb.extract_color_features(crops), tr.color_features(crops))
    live = hyb.extract_deep_features(crops)
    mine = tr.deep_features(data["trunk"], crops)
    assert np.abs(live - mine).max() <= 1e-5 * np.abs(live).max()

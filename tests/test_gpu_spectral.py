"""Opt-in device spectral clustering inside HybridTeamClassifier.fit: same labels as the
reference's sklearn solver wherever the clustering is well defined, predictions untouched, and much faster."""
import os
import sys
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import golden_crops  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "team_reference.npz"))


@pytest.fixture(scope="module")
def data():
    _, crops, labels, _, tids = golden_crops()
    n_fit = int((labels >= 0).sum())
    return crops[:n_fit], labels[:n_fit], tids[:n_fit]


def test_device_solver_gives_sklearn_labels_with_a_sane_bandwidth(ctx, data):
    from hvb import HybridTeamClassifier
    from hvb.models import build_trunk
    crops, truth, _ = data
    trunk = build_trunk(0)
    out = {}
    for mode in ("sklearn", "device"):
        clf = HybridTeamClassifier(device="cuda:0", trunk=trunk, spectral=mode, affinity_gamma="scale")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        clf.fit(crops)
        out[mode] = (clf.cluster_labels.copy(), time.perf_counter() - t0, clf.affinity_matrix_)
    assert np.array_equal(out["sklearn"][2], out["device"][2])
    off = out["device"][2][~np.eye(len(crops), dtype=bool)]
    assert off.max() > 0.1                                              # gamma = 1/625 does not underflow
    assert np.array_equal(out["device"][0], out["sklearn"][0])
    lab = out["device"][0]
    assert np.array_equal(lab, truth) or np.array_equal(lab, 1 - truth)   # the two jersey colours
    print("fit N=%d: sklearn solver %.3f s, device solver %.3f s" % (len(crops), out["sklearn"][1], out["device"][1]))


def test_reference_bandwidth_predictions_do_not_depend_on_the_solver(ctx, data):
    """gamma = 1 (the reference): the affinity is numerically the identity and the labels are arbitrary with either
    solver, but predict() never reads them (team_hybrid.py:241-262) — the golden predictions must come out."""
    from hvb import HybridTeamClassifier
    from hvb.models import build_trunk
    crops, _, tids = data
    clf = HybridTeamClassifier(device="cuda:0", trunk=build_trunk(0), affinity_mode=1, spectral="device")
    t0 = time.perf_counter()
    clf.fit(crops)
    dt = time.perf_counter() - t0
    assert set(np.unique(clf.cluster_labels)) <= {0, 1} and len(clf.cluster_labels) == len(crops)
    assert clf.clusterer.affinity_matrix_ is clf.affinity_matrix_ and (np.diag(clf.affinity_matrix_) == 1).all()
    per = len(crops) // 4
    preds = np.concatenate([clf.predict(crops[f * per:(f + 1) * per], tids[f * per:(f + 1) * per]) for f in range(4)])
    assert np.array_equal(preds, GOLD["predict"])
    print("fit N=%d at gamma=1 with the device solver: %.3f s" % (len(crops), dt))
    with pytest.raises(ValueError):
        HybridTeamClassifier(device="cuda:0", trunk=clf.feature_extractor, spectral="lobpcg")


def test_gathered_fit_size(ctx):
    """N = 2000 feature rows (8 clips gathered): the dense solve stays interactive."""
    from hvb import HybridTeamClassifier
    from hvb.models import build_trunk
    rng = np.random.default_rng(0)
    centres = rng.normal(0, 1, (2, 625))
    which = rng.integers(0, 2, 2000)
    feats = torch.from_numpy(centres[which] * 3 + rng.normal(0, 1, (2000, 625))).cuda()
    clf = HybridTeamClassifier(device="cuda:0", trunk=build_trunk(0), spectral="device", affinity_gamma="scale")
    clf.fit_features(feats)                                            # warm-up (cuSOLVER handles, kernels)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    clf.fit_features(feats)
    dt = time.perf_counter() - t0
    lab = clf.cluster_labels
    assert np.array_equal(lab, which) or np.array_equal(lab, 1 - which)
    print("fit_features N=2000 with the device solver: %.3f s" % dt)
    assert dt < 5.0

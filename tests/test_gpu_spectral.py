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


# ---------------------------------------------------------------- K8 kernels against their numpy twin (tests/spectral_twin.py)
def _affinity(seed, n, d=20):
    rng = np.random.default_rng(seed)
    x = np.vstack([rng.normal(0, 1, (n // 2, d)), rng.normal(0, 1, (n - n // 2, d)) + 2.5])
    x = x[rng.permutation(n)]
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    return np.exp(-d2 / d)


@pytest.mark.parametrize("n", [33, 257, 1000])
def test_k8_laplacian_matvec_gram_rotate_match_the_twin(ctx, n):
    from spectral_twin import NumpyOps
    a = _affinity(n, n)
    m, dd = ctx.laplacian_normalize(torch.from_numpy(a).cuda())
    m_ref, dd_ref = NumpyOps.normalize(a)
    assert np.allclose(dd.cpu().numpy(), dd_ref, rtol=1e-14, atol=0)
    assert np.allclose(m.cpu().numpy(), m_ref, rtol=1e-13, atol=1e-300)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((8, n))
    xd = torch.from_numpy(x).cuda()
    y = ctx.sym_block_matvec(m, xd, 0.25)
    y_ref = NumpyOps.matvec(m_ref, x, 0.25)
    assert np.abs(y.cpu().numpy() - y_ref).max() <= 1e-12 * np.abs(y_ref).max()
    g = ctx.block_gram(xd, y, 0).cpu().numpy()
    assert np.abs(g[:64] - NumpyOps.gram(x, y_ref, 0)[:64]).max() <= 1e-11 * np.abs(g[:64]).max()
    rinv = ctx.block_gram(xd, xd, 1)
    assert rinv[64].item() == 0.0
    ctx.block_rotate(xd, None, rinv[:64])
    q = xd.cpu().numpy()
    assert np.abs(q @ q.T - np.eye(8)).max() <= 1e-12 * max(1.0, np.linalg.cond(x @ x.T))     # orthonormal after one Cholesky QR
    lam = rng.standard_normal(8)
    rot = np.linalg.qr(rng.standard_normal((8, 8)))[0]
    x2, y2 = q.copy(), y_ref.copy()
    res_ref = NumpyOps.rotate(x2, y2, rot, lam)
    yd = torch.from_numpy(y_ref).cuda()
    res = ctx.block_rotate(xd, yd, torch.from_numpy(np.ascontiguousarray(rot)).cuda(), torch.from_numpy(lam).cuda())
    assert np.abs(xd.cpu().numpy() - x2).max() <= 1e-13 and np.abs(yd.cpu().numpy() - y2).max() <= 1e-12 * np.abs(y2).max()
    assert np.allclose(res.cpu().numpy(), res_ref, rtol=1e-10)
    # a collapsed block is reported, not silently factorised
    bad = torch.from_numpy(np.repeat(x[:1], 8, axis=0).copy()).cuda()
    assert ctx.block_gram(bad, bad, 1)[64].item() == 1.0


@pytest.mark.parametrize("n,k", [(135, 2), (135, 3), (2000, 2)])
def test_k8_kmeans_lloyd_matches_the_twin_and_sklearn(ctx, n, k):
    from sklearn.cluster import KMeans
    from spectral_twin import NumpyOps
    from hvb.spectral import _DeviceOps, kmeans_best_of
    rng = np.random.default_rng(n + k)
    cen = rng.normal(0, 2.0, (3, 2))
    emb = np.vstack([rng.normal(0, 0.4, (n // 3, 2)) + cen[i] for i in range(3)] + [rng.normal(0, 0.4, (n - 3 * (n // 3), 2)) + cen[0]])
    x = emb - emb.mean(0)
    init = np.stack([x[rng.choice(len(x), k, replace=False)] for _ in range(10)])
    tol = float(np.mean(np.var(x, axis=0)) * 1e-4)
    got = [t.cpu().numpy() for t in ctx.kmeans_lloyd(torch.from_numpy(x).cuda(), torch.from_numpy(init).cuda(), 300, tol)]
    ref = NumpyOps.kmeans(x, init, 300, tol)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[3], ref[3]) and np.array_equal(got[4], ref[4])
    assert np.allclose(got[1], ref[1], rtol=1e-12, atol=1e-14) and np.allclose(got[2], ref[2], rtol=1e-12)
    km = KMeans(n_clusters=k, n_init=10, random_state=np.random.RandomState(3)).fit(emb)
    lab = kmeans_best_of(emb, k, 10, np.random.RandomState(3), _DeviceOps(ctx))
    assert np.array_equal(lab, km.labels_)


@pytest.mark.parametrize("n", [220, 2000])
def test_k8_subspace_embedding_equals_the_dense_solver(ctx, n):
    from hvb.spectral import _DeviceOps, spectral_embedding_dense, spectral_embedding_subspace
    a = torch.from_numpy(_affinity(7, n)).cuda()
    info = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    got = spectral_embedding_subspace(a, 2, _DeviceOps(ctx), info=info)
    t1 = time.perf_counter()
    dense = spectral_embedding_dense(a, 2).cpu().numpy()
    t2 = time.perf_counter()
    assert np.abs(got - dense).max() <= 1e-8 * np.abs(dense).max()
    print("N=%d: subspace solver %.1f ms (%d outer rounds, %d matvecs, residual %.1e), cuSOLVER eigh %.1f ms"
          % (n, 1e3 * (t1 - t0), info["outer_iterations"], info["matvecs"], info["residuals"].max(), 1e3 * (t2 - t1)))

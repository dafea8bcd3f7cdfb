"""K4a (float64 path) and K4b parity vs scikit-learn / the restated supervision matching costs."""
import numpy as np
import pytest
import torch

from oracle import supervision_restated as svr
from oracle import team_reference as tr

pytestmark = pytest.mark.gpu


def feature_matrix(seed, n, d=625, dup=6):
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 1, (n, d))
    for k in range(min(dup, n // 2)):               # near-duplicate rows so the affinity is non-trivial
        x[n - 1 - k] = x[k] + rng.normal(0, 0.02, d)
    x[:, 5] = 3.0                                   # constant column -> scale 1
    x[:, 6] *= 1e-9                                 # tiny-variance column (like random-init deep features)
    return x


@pytest.mark.parametrize("n", [4, 40, 250])
def test_standardize_matches_sklearn(ctx, n):
    x = feature_matrix(n, n)
    sc, ref = tr.standardize_fit(x)
    mean, scale, xs = ctx.standardize(torch.from_numpy(x).cuda())
    np.testing.assert_allclose(mean.cpu().numpy(), sc.mean_, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(scale.cpu().numpy(), sc.scale_, rtol=1e-9)
    assert scale.cpu().numpy()[5] == 1.0
    np.testing.assert_allclose(xs.cpu().numpy(), ref, rtol=1e-9, atol=1e-9)
    xt = ctx.scale_transform(torch.from_numpy(x[:3].copy()).cuda(), mean, scale).cpu().numpy()
    np.testing.assert_allclose(xt, sc.transform(x[:3]), rtol=1e-9, atol=1e-9)


def check_affinity(d2, a, x, gamma=1.0):
    d2_ref, a_ref = tr.rbf_affinity(x, gamma)
    n = len(x)
    assert (np.diag(d2) == 0).all() and (np.diag(a) == 1.0).all()
    off = ~np.eye(n, dtype=bool)
    assert np.abs(d2[off] - d2_ref[off]).max() <= 1e-3 * np.abs(d2_ref[off]).max()
    assert (np.abs(d2 - d2_ref)[off] <= 1e-3 * d2_ref[off] + 1e-9).all()
    big = a_ref > 1e-300
    assert big.sum() > n                            # fixture has non-trivial off-diagonal affinities
    assert (np.abs(a[big] - a_ref[big]) <= 1e-3 * a_ref[big]).all()
    assert (np.abs(a[~big] - a_ref[~big]) <= 1e-300).all()
    assert np.array_equal(a, a.T)


@pytest.mark.parametrize("n,d", [(4, 625), (40, 625), (250, 627), (333, 100)])
def test_affinity_fp64_mode(ctx, n, d):
    x = tr.standardize_fit(feature_matrix(n + d, n, d, dup=min(6, n // 2)))[1]
    d2, a = ctx.gram_affinity(torch.from_numpy(x).cuda(), 1.0, mode=1)
    check_affinity(d2.cpu().numpy(), a.cpu().numpy(), x)
    d2h, ah = ctx.gram_affinity_host(x, 1.0, mode=1)
    assert np.array_equal(d2h, d2.cpu().numpy()) and np.array_equal(ah, a.cpu().numpy())


@pytest.mark.parametrize("na,nb", [(0, 5), (5, 0), (1, 1), (12, 11), (40, 37), (64, 64), (300, 257)])
def test_iou_cost_bit_exact(ctx, na, nb):
    rng = np.random.default_rng(na * 100 + nb)
    from hvb.synth import random_boxes
    a64 = random_boxes(rng, na, 300, 300, 20, 150, dtype=np.float64)[0]          # Kalman track boxes (float64)
    b32, scores = random_boxes(rng, nb, 300, 300, 20, 150, dtype=np.float32)     # detections (float32)
    if na > 3 and nb > 3:
        a64[0] = [5, 5, 5, 5]; b32[0] = [5, 5, 5, 5]                             # zero-area pair -> nan_to_num -> IoU 0
        a64[1] = b32[1]                                                          # identical boxes
    a32 = random_boxes(rng, na, 300, 300, 20, 150, dtype=np.float32)[0]
    b64 = random_boxes(rng, nb, 300, 300, 20, 150, dtype=np.float64)[0]
    sc = np.linspace(0.3, 0.9, nb)
    for a, b in ((a64, b32), (a64, b64), (a32, b64)):      # ByteTrack never pairs two float32 sides
        ref = svr.iou_distance(a, b)
        got = ctx.iou_cost_host(a, b)
        assert got.shape == ref.shape and np.array_equal(got, ref)
        if na and nb:
            assert np.array_equal(ctx.iou_cost_host(a, b, sc), svr.fuse_score(ref.copy(), sc))


def test_iou_cost_batched_problems(ctx):
    rng = np.random.default_rng(0)
    from hvb.synth import random_boxes
    sizes = [(3, 4), (0, 2), (10, 7), (5, 0), (33, 40)]
    A = [random_boxes(rng, na, 200, 200, 20, 90, np.float64)[0] for na, _ in sizes]
    B = [random_boxes(rng, nb, 200, 200, 20, 90, np.float64)[0] for _, nb in sizes]
    a_off = np.cumsum([0] + [s[0] for s in sizes]).astype(np.int32)
    b_off = np.cumsum([0] + [s[1] for s in sizes]).astype(np.int32)
    o_off = np.cumsum([0] + [s[0] * s[1] for s in sizes]).astype(np.int64)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    out = ctx.iou_cost(dev(np.concatenate(A)), dev(np.concatenate(B)), None, dev(a_off), dev(b_off), dev(o_off[:-1].copy()),
                       len(sizes), 33, 40, int(o_off[-1])).cpu().numpy()
    for p, (na, nb) in enumerate(sizes):
        assert np.array_equal(out[o_off[p]:o_off[p + 1]].reshape(na, nb), svr.iou_distance(A[p], B[p]))

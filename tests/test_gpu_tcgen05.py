"""K4a tensor-core path: the tcgen05 (split-TF32, TMEM-accumulated) Gram and the mode-0 affinity.
Kept in its own file so it can be run in a separate process from the rest of the GPU suite."""
import numpy as np
import pytest
import torch

from oracle import team_reference as tr
from test_gpu_affinity import check_affinity, feature_matrix

pytestmark = [pytest.mark.gpu, pytest.mark.tcgen05]


@pytest.mark.parametrize("n,d", [(128, 32), (4, 625), (40, 625), (250, 625), (250, 627), (300, 100), (1000, 625)])
def test_gram_tcgen05_matches_float64(ctx, n, d):
    x = tr.standardize_fit(feature_matrix(n + d, n, d, dup=min(6, n // 2)))[1]
    g = ctx.gram_tc(torch.from_numpy(x).cuda()).cpu().numpy().astype(np.float64)
    ref = x @ x.T
    # split-TF32 (hi.hi + hi.lo + lo.hi), fp32 accumulate: ~1e-6 of |x||y|
    tol = 2e-5 * np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    assert (np.abs(g - ref) <= tol + 1e-6).all(), float((np.abs(g - ref) / (tol + 1e-6)).max())


@pytest.mark.parametrize("n,d", [(40, 625), (250, 627), (1000, 625)])
def test_affinity_tensor_core_mode(ctx, n, d):
    x = tr.standardize_fit(feature_matrix(n + d, n, d, dup=min(6, n // 2)))[1]
    d2, a = ctx.gram_affinity(torch.from_numpy(x).cuda(), 1.0, mode=0)
    check_affinity(d2.cpu().numpy(), a.cpu().numpy(), x)


def test_golden_reference_affinity(ctx):
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "team_reference.npz"))
    xs = gold["features_scaled"]
    _, a = ctx.gram_affinity(torch.from_numpy(np.ascontiguousarray(xs)).cuda(), 1.0, mode=0)
    a = a.cpu().numpy()
    ref = gold["affinity"]
    big = ref > 1e-300
    assert (np.abs(a[big] - ref[big]) <= 1e-3 * ref[big]).all()
    assert (np.abs(a[~big] - ref[~big]) <= 1e-300).all()

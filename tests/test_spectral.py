"""CPU: hvb/spectral.py (opt-in device spectral clustering) against scikit-learn's
spectral_embedding / SpectralClustering / KMeans on affinities where the answer is well defined (distinct leading
eigenvalues, separated clusters).  The dense cross-check solver (torch.linalg.eigh) runs on CPU tensors here; the
subspace solver's and the k-means driver's HOST logic run over the numpy twin of the K8 kernels (tests/spectral_twin.py) —
the kernels themselves are checked against that twin on the GPU (tests/test_gpu_spectral.py)."""
import warnings

import numpy as np
import pytest
import torch

from hvb.spectral import DeviceSpectralClustering, kmeans_best_of, spectral_embedding_dense, spectral_embedding_subspace
from spectral_twin import NumpyOps


def _blobs(seed, n_per=(60, 45), d=20, sep=4.0):
    rng = np.random.default_rng(seed)
    x = np.vstack([rng.normal(0, 1, (n_per[0], d)), rng.normal(0, 1, (n_per[1], d)) + sep / np.sqrt(d) * 3])
    truth = np.repeat([0, 1], n_per)
    perm = rng.permutation(len(x))
    x, truth = x[perm], truth[perm]
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    return np.exp(-d2 / d), truth


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_embedding_matches_sklearn(seed):
    from sklearn.manifold import spectral_embedding
    a, _ = _blobs(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = spectral_embedding(a, n_components=2, eigen_solver="arpack", random_state=np.random.RandomState(42), drop_first=False)
    got = spectral_embedding_dense(torch.from_numpy(a), 2).numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_labels_match_sklearn_spectral_clustering(seed):
    from sklearn.cluster import SpectralClustering
    a, truth = _blobs(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = SpectralClustering(n_clusters=2, affinity="precomputed", n_init=10, random_state=42).fit_predict(a)
    got = DeviceSpectralClustering(2, 10, 42, solver="eigh", kmeans="sklearn").fit_predict(torch.from_numpy(a))
    assert np.array_equal(got, ref)                                   # same RNG stream -> same numbering, not only the same partition
    assert np.array_equal(got, truth) or np.array_equal(got, 1 - truth)


def test_isolated_nodes_and_three_clusters():
    rng = np.random.default_rng(5)
    a, _ = _blobs(5, (30, 30))
    a = np.pad(a, ((0, 1), (0, 1)))                                     # one node connected to nothing
    a[-1, -1] = 1.0
    emb = spectral_embedding_dense(torch.from_numpy(a), 3).numpy()
    assert np.isfinite(emb).all()
    x = np.vstack([rng.normal(c, 0.3, (25, 4)) for c in (0, 3, 6)])
    d2 = ((x[:, None] - x[None]) ** 2).sum(-1)
    lab = DeviceSpectralClustering(3, solver="eigh", kmeans="sklearn").fit_predict(torch.from_numpy(np.exp(-d2)))
    assert [len(set(lab[i * 25:(i + 1) * 25])) for i in range(3)] == [1, 1, 1] and len(set(lab)) == 3


# ---------------------------------------------------------------- subspace solver + device k-means: host logic over the twin
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_subspace_embedding_matches_dense_and_sklearn(seed):
    from sklearn.manifold import spectral_embedding
    a, _ = _blobs(seed)
    info = {}
    got = spectral_embedding_subspace(a, 2, NumpyOps(), info=info)
    dense = spectral_embedding_dense(torch.from_numpy(a), 2).numpy()
    assert np.abs(got - dense).max() <= 1e-8 * max(1.0, np.abs(dense).max())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = spectral_embedding(a, n_components=2, eigen_solver="arpack", random_state=np.random.RandomState(42), drop_first=False)
    assert np.abs(got - ref).max() <= 1e-7 * max(1.0, np.abs(ref).max())
    assert info["outer_iterations"] <= 40 and info["residuals"].max() <= 1e-9


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_kmeans_driver_equals_sklearn_kmeans(seed):
    """Same labels (numbering included) and the same best inertia as KMeans.fit_predict with the same random stream."""
    from sklearn.cluster import KMeans
    rng = np.random.default_rng(seed)
    emb = np.vstack([rng.normal(0, 0.4, (70, 2)) + [1.5, 0], rng.normal(0, 0.5, (50, 2)) - [1.0, 0.5], rng.normal(0, 0.3, (15, 2)) + [0, 2.0]])
    for k in (2, 3):
        ref = KMeans(n_clusters=k, n_init=10, random_state=np.random.RandomState(7)).fit(emb)
        info = {}
        got = kmeans_best_of(emb, k, 10, np.random.RandomState(7), NumpyOps(), info=info)
        assert np.array_equal(got, ref.labels_)
        assert abs(info["kmeans_inertia"][info["kmeans_best"]] - ref.inertia_) <= 1e-9 * ref.inertia_


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_full_device_path_labels_match_sklearn(seed):
    from sklearn.cluster import SpectralClustering
    a, truth = _blobs(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = SpectralClustering(n_clusters=2, affinity="precomputed", n_init=10, random_state=42).fit_predict(a)
    got = DeviceSpectralClustering(2, 10, 42, ops=NumpyOps()).fit_predict(a)
    assert np.array_equal(got, ref)


def test_subspace_solver_reports_rank_loss():
    a = np.ones((40, 40))                                           # rank one: an 8-vector block cannot stay independent
    with pytest.raises(FloatingPointError):
        spectral_embedding_subspace(a, 2, NumpyOps())

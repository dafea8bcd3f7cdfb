"""Opt-in dense spectral clustering of a precomputed affinity on the GPU.

The reference clusters with ``SpectralClustering(affinity='rbf', gamma=1.0)`` (team_hybrid.py:185-193); sklearn's
default ARPACK shift-invert solve of the N x N normalised Laplacian takes seconds at N = 220 (the affinity of 625-d
standardised features under gamma = 1 is numerically the identity, the worst case for an iterative solver) and the
reference's ``predict`` never reads the labels — they only feed the printed cluster statistics and the 0/1 ordering.
This module restates ``sklearn.manifold.spectral_embedding(norm_laplacian=True, drop_first=False)`` + ``KMeans`` on
the device the affinity already lives on, with libhvb's own kernels (K8, csrc/k8_spectral.cu):

  solver="subspace" (default)  normalised Laplacian pieces (hvb_laplacian_normalize), then block subspace iteration on
      M = D^-1/2 A0 D^-1/2 with 8 vectors: a few power steps Y = (M + shift I) X (hvb_sym_block_matvec, one HBM pass over
      the N x N matrix each), Cholesky-QR re-orthonormalisation (hvb_block_gram / hvb_block_rotate), Rayleigh-Ritz on
      the 8 x 8 projected matrix (numpy on the host: 64 numbers) and a residual test |M v - lambda v| <= tol per wanted
      pair.  The host only steers; every pass over N-sized data is a kernel.
  solver="eigh"                the dense cross-check: ``torch.linalg.eigh`` (cuSOLVER syevd) on the same Laplacian.
  kmeans="device" (default)    k-means++ seeding by sklearn's own routine with the reference's random stream (host, N x k
      numbers), then every Lloyd run of the n_init initialisations in ONE launch (hvb_kmeans_lloyd, one CTA each), best
      inertia picked like KMeans.fit does.
  kmeans="sklearn"             sklearn's KMeans on the host.

Converged eigenvectors instead of ARPACK's: identical labels whenever the clusters are separated, arbitrary (as with
ARPACK) inside a degenerate eigenspace — hence opt-in: ``HybridTeamClassifier(spectral="device")``.
"""
from __future__ import annotations

import numpy as np
import torch


def spectral_embedding_dense(affinity: torch.Tensor, n_components: int) -> torch.Tensor:
    """float64[N,N] symmetric affinity -> float64[N, n_components]; follows sklearn/manifold/_spectral_embedding.py
    (scipy ``csgraph.laplacian(normed=True)``, diagonal forced to 1, eigenvectors of the smallest eigenvalues of L,
    divided by sqrt(degree), deterministic sign flip)."""
    a = affinity.to(torch.float64).clone()
    a.fill_diagonal_(0.0)
    w = a.sum(0)
    isolated = w == 0
    dd = torch.where(isolated, torch.ones_like(w), torch.sqrt(w))
    lap = -(a / dd[None, :]) / dd[:, None]
    lap.fill_diagonal_(1.0)
    _, vec = torch.linalg.eigh(lap)                       # ascending eigenvalues of L
    emb = vec[:, :n_components].T / dd[None, :]           # rows = vectors, as sklearn's `embedding`
    idx = emb.abs().argmax(1)
    sign = torch.sign(emb[torch.arange(emb.shape[0], device=emb.device), idx])
    sign = torch.where(sign == 0, torch.ones_like(sign), sign)
    return (emb * sign[:, None]).T.contiguous()


class _DeviceOps:
    """The K8 entry points behind the solver (hvb.runtime.Context); tests drive the same host logic with a numpy twin."""

    def __init__(self, ctx):
        self.ctx = ctx

    def to_dev(self, a: np.ndarray) -> torch.Tensor:
        return self.ctx.to_device(np.ascontiguousarray(a, np.float64))

    @staticmethod
    def to_host(t: torch.Tensor) -> np.ndarray:
        return t.cpu().numpy()

    @staticmethod
    def concat(parts):
        return torch.cat(list(parts))

    def normalize(self, a):
        return self.ctx.laplacian_normalize(a)

    def matvec(self, m, x, shift, out=None):
        return self.ctx.sym_block_matvec(m, x, shift, out)

    def gram(self, a, b, mode):
        return self.ctx.block_gram(a, b, mode)

    def rotate(self, x, y, q, lam=None):
        return self.ctx.block_rotate(x, y, q, lam)

    def kmeans(self, x, init, max_iter, tol):
        return self.ctx.kmeans_lloyd(x, init, max_iter, tol)


BLOCK = 8            # hvb_spectral_block(): vectors per block
SMALL_N = 32         # below this the 8-vector block is (nearly) the whole space: the dense solver handles it


def _sign_flip(emb: np.ndarray) -> np.ndarray:
    """sklearn.utils.extmath._deterministic_vector_sign_flip on rows."""
    idx = np.abs(emb).argmax(1)
    sign = np.sign(emb[np.arange(emb.shape[0]), idx])
    sign[sign == 0] = 1
    return emb * sign[:, None]


def spectral_embedding_subspace(affinity, n_components: int, ops, tol: float = 1e-10, power_steps: int = 3,
                                max_outer: int = 300, seed: int = 0, info: dict = None) -> np.ndarray:
    """float64[N,N] affinity (device tensor for the device ops) -> float64[N, n_components] embedding on the host, the
    same quantity as spectral_embedding_dense.  `ops`: _DeviceOps (product) or a twin with the same methods (tests)."""
    n = int(affinity.shape[0])
    if not 1 <= n_components <= BLOCK - 2:
        raise ValueError("n_components must be in [1, %d]" % (BLOCK - 2))
    m, dd = ops.normalize(affinity)
    dd_h = ops.to_host(dd)
    # A (an RBF Gram matrix) is PSD, so M = D^-1/2 A D^-1/2 - D^-1 has eigenvalues >= -max(1/degree): with this shift
    # the iteration matrix is PSD and the wanted (largest) eigenvalues of M are also the largest in magnitude
    shift = float(min(1.0, (1.0 / (dd_h * dd_h)).max()))
    rng = np.random.RandomState(seed)
    x0 = rng.standard_normal((BLOCK, n))
    x0[0] = dd_h                                            # D^1/2 1 is the exact top eigenvector of a connected graph
    x = ops.to_dev(x0)
    y = ops.to_dev(np.zeros((BLOCK, n)))
    flags = []

    def orthonormalize():
        for _ in range(2):                                  # Cholesky QR twice: orthogonal to machine precision
            rinv = ops.gram(x, x, 1)
            flags.append(rinv[64:65])
            ops.rotate(x, None, rinv[:64])

    orthonormalize()
    lam = np.zeros(BLOCK)
    res = np.full(n_components, np.inf)
    outer = 0
    for outer in range(1, max_outer + 1):
        for _ in range(power_steps):
            ops.matvec(m, x, shift, out=y)
            x, y = y, x
        orthonormalize()
        ops.matvec(m, x, shift, out=y)
        small = ops.to_host(ops.concat([ops.gram(x, y, 0)[:64]] + flags))          # the one read of this round
        if small[64:].any():
            raise FloatingPointError("subspace iteration: the vector block lost rank (affinity with fewer than %d "
                                     "independent directions?)" % BLOCK)
        flags.clear()
        h = small[:64].reshape(BLOCK, BLOCK)
        w, q = np.linalg.eigh((h + h.T) * 0.5)
        order = np.argsort(-w, kind="stable")
        lam, q = w[order], np.ascontiguousarray(q[:, order])
        res2 = ops.to_host(ops.rotate(x, y, ops.to_dev(q), ops.to_dev(lam)))
        res = np.sqrt(np.maximum(res2[:n_components], 0.0))
        if res.max() <= tol * max(abs(lam[0]), 1e-300):
            break
    if info is not None:
        info.update(outer_iterations=outer, matvecs=outer * (power_steps + 1), residuals=res, eigenvalues=lam - shift, shift=shift)
    vec = ops.to_host(x)[:n_components]                    # rows = Ritz vectors, largest eigenvalue of M first
    return np.ascontiguousarray(_sign_flip(vec / dd_h[None, :]).T)


def _same_clustering(a: np.ndarray, b: np.ndarray, k: int) -> bool:
    """sklearn.cluster._k_means_common._is_same_clustering: equal up to a permutation of the labels."""
    mapping = np.full(k, -1, np.int64)
    for i, j in zip(a, b):
        if mapping[i] == -1:
            mapping[i] = j
        elif mapping[i] != j:
            return False
    return True


def kmeans_best_of(emb: np.ndarray, n_clusters: int, n_init: int, rs: np.random.RandomState, ops, max_iter: int = 300,
                   tol: float = 1e-4, info: dict = None) -> np.ndarray:
    """KMeans(n_clusters, n_init=n_init, random_state=rs).fit_predict(emb) with the Lloyd runs on the device: sklearn's
    own preamble (centring, tolerance, k-means++ seeding with the caller's random stream, sklearn/cluster/_kmeans.py
    KMeans.fit), one hvb_kmeans_lloyd launch for all initialisations, then KMeans.fit's best-inertia selection.  A run
    that empties a cluster (sklearn relocates the farthest points then) hands the whole call to sklearn."""
    from sklearn.cluster import KMeans, kmeans_plusplus
    x = np.array(emb, dtype=np.float64, order="C")
    state = rs.get_state()
    tol_abs = float(np.mean(np.var(x, axis=0)) * tol)
    x -= x.mean(axis=0)
    xsn = (x * x).sum(1)
    inits = np.stack([kmeans_plusplus(x, n_clusters, x_squared_norms=xsn, random_state=rs)[0] for _ in range(n_init)])
    labels, _centers, inertia, n_iter, flags = (ops.to_host(t) for t in ops.kmeans(ops.to_dev(x), ops.to_dev(inits), max_iter, tol_abs))
    if flags.any():
        rs.set_state(state)
        return KMeans(n_clusters=n_clusters, n_init=n_init, random_state=rs).fit_predict(emb)
    best = 0
    for i in range(1, n_init):
        if inertia[i] < inertia[best] and not _same_clustering(labels[i], labels[best], n_clusters):
            best = i
    if info is not None:
        info.update(kmeans_best=best, kmeans_inertia=inertia, kmeans_iterations=n_iter)
    return labels[best].astype(np.int32)


class DeviceSpectralClustering:
    """The attributes the reference reads from its sklearn clusterer: ``affinity_matrix_`` (set by the caller),
    ``labels_``, ``fit_predict``."""

    def __init__(self, n_clusters: int = 2, n_init: int = 10, random_state: int = 42, solver: str = "subspace",
                 kmeans: str = "device", ops=None):
        if solver not in ("subspace", "eigh") or kmeans not in ("device", "sklearn"):
            raise ValueError("solver: 'subspace' | 'eigh'; kmeans: 'device' | 'sklearn'")
        self.n_clusters, self.n_init, self.random_state = n_clusters, n_init, random_state
        self.solver, self.kmeans, self.ops = solver, kmeans, ops
        self.affinity_matrix_ = None
        self.labels_ = None
        self.embedding_ = None
        self.info_ = {}

    def _ops(self, affinity):
        if self.ops is None:
            from .runtime import get_context
            self.ops = _DeviceOps(get_context(affinity.device))      # raises without a B200: no CPU path in the product
        return self.ops

    def fit_predict(self, affinity: torch.Tensor) -> np.ndarray:
        if self.ops is None and not isinstance(affinity, torch.Tensor):
            from .runtime import get_context
            affinity = get_context("cuda:0").to_device(np.asarray(affinity, np.float64))     # loud without a B200
        n = affinity.shape[0]
        emb = None
        if self.solver == "subspace" and n >= SMALL_N and self.n_clusters <= BLOCK - 2:
            try:
                emb = spectral_embedding_subspace(affinity, self.n_clusters, self._ops(affinity), info=self.info_)
            except FloatingPointError:
                # fewer than 8 independent directions (e.g. the reference's gamma = 1 affinity, numerically the identity:
                # every node isolated, M = 0): any basis of that eigenspace is as good as another, take the dense solver's
                self.info_["subspace_rank_loss"] = True
        if emb is None:
            aff = affinity if isinstance(affinity, torch.Tensor) else torch.from_numpy(np.asarray(affinity))
            emb = spectral_embedding_dense(aff, self.n_clusters).cpu().numpy()
        rs = np.random.RandomState(self.random_state)
        rs.uniform(-1, 1, emb.shape[0])                   # sklearn draws ARPACK's start vector from the same stream first
        self.embedding_ = emb
        if self.kmeans == "device" and self.n_clusters <= 8:
            self.labels_ = kmeans_best_of(emb, self.n_clusters, self.n_init, rs, self._ops(affinity), info=self.info_)
        else:
            from sklearn.cluster import KMeans
            self.labels_ = KMeans(n_clusters=self.n_clusters, n_init=self.n_init, random_state=rs).fit_predict(emb)
        return self.labels_

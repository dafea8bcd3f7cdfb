"""numpy twin of the K8 entry points (csrc/k8_spectral.cu) with the same layouts and flags — TEST INFRASTRUCTURE: lets
the CPU suite drive hvb/spectral.py's host logic (iteration control, Rayleigh-Ritz, k-means seeding and selection); the
GPU suite checks the kernels themselves against these functions."""
import numpy as np


class NumpyOps:
    @staticmethod
    def to_dev(a):
        return np.array(a, dtype=np.float64, order="C")

    @staticmethod
    def to_host(a):
        return np.asarray(a)

    @staticmethod
    def concat(parts):
        return np.concatenate([np.asarray(p).reshape(-1) for p in parts])

    @staticmethod
    def normalize(a):
        a = np.array(a, dtype=np.float64)
        np.fill_diagonal(a, 0.0)
        w = a.sum(axis=0)
        dd = np.where(w == 0, 1.0, np.sqrt(w))
        return a / dd[None, :] / dd[:, None], dd

    @staticmethod
    def matvec(m, x, shift, out=None):
        y = (m @ x.T).T + shift * x
        if out is not None:
            out[...] = y
            return out
        return y

    @staticmethod
    def gram(a, b, mode):
        g = a @ b.T
        out = np.zeros(65)
        if mode == 0:
            out[:64] = g.reshape(-1)
            return out
        try:
            r = np.linalg.cholesky(g).T                     # upper: g = r.T @ r
            out[:64] = np.linalg.inv(r).reshape(-1)
        except np.linalg.LinAlgError:
            out[:64] = np.eye(8).reshape(-1)
            out[64] = 1.0
        return out

    @staticmethod
    def rotate(x, y, q, lam=None):
        q = np.asarray(q).reshape(8, 8)
        x[...] = (x.T @ q).T
        if y is not None:
            y[...] = (y.T @ q).T
        if lam is None:
            return None
        d = y - np.asarray(lam)[:, None] * x
        return (d * d).sum(1)

    @staticmethod
    def kmeans(x, init, max_iter, tol):
        """sklearn _kmeans_single_lloyd per initialisation, as kmeans_lloyd_kernel runs it."""
        n_init, k, d = init.shape
        n = x.shape[0]
        labels_all = np.zeros((n_init, n), np.int32)
        centers_all = np.zeros((n_init, k, d))
        inertia, n_iter, flags = np.zeros(n_init), np.zeros(n_init, np.int32), np.zeros(n_init, np.int32)
        for r in range(n_init):
            cen = init[r].copy()
            labels = np.full(n, -1)
            strict = False
            it = 0
            for it in range(1, max_iter + 1):
                new = np.argmin((cen * cen).sum(1)[None, :] - 2.0 * x @ cen.T, axis=1)
                changed = not np.array_equal(new, labels)
                labels = new
                cnew = cen.copy()
                for c in range(k):
                    sel = labels == c
                    if sel.any():
                        cnew[c] = x[sel].sum(0) / sel.sum()
                    else:
                        flags[r] = 1
                shift = ((cnew - cen) ** 2).sum()
                cen = cnew
                if flags[r]:
                    break
                if not changed:
                    strict = True
                    break
                if shift <= tol:
                    break
            if not strict and not flags[r]:
                labels = np.argmin((cen * cen).sum(1)[None, :] - 2.0 * x @ cen.T, axis=1)
            labels_all[r], centers_all[r] = labels, cen
            inertia[r] = ((x - cen[labels]) ** 2).sum()
            n_iter[r] = it
        return labels_all, centers_all, inertia, n_iter, flags

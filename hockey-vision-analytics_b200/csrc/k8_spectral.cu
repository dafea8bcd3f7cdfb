// K8 — device spectral clustering of the precomputed affinity: what
// SpectralClustering(affinity='precomputed').fit_predict does after K4a in HybridTeamClassifier.fit
// (hockey/common/team_hybrid.py:185-193): sklearn.manifold.spectral_embedding(norm_laplacian=True, drop_first=False)
// followed by KMeans(n_init=10) on the N x k embedding.  Opt-in (it changes which solver produces the labels).
//
// Pieces (all float64, all stream-ordered, no host synchronisation inside):
//   degree_kernel / normalize_kernel   scipy.sparse.csgraph.laplacian(normed=True) on a dense matrix: zero the diagonal,
//                                      w = column sums, dd = sqrt(w) (1 for isolated nodes), M = D^-1/2 A0 D^-1/2
//                                      (L = I - M; the embedding is the top eigenvectors of M)
//   sym_block_matvec_kernel            Y = (M + shift I) X for a block of 8 vectors: the one pass over the N x N matrix
//                                      per subspace-iteration step (HBM-bound: N*N*8 bytes, 8 DFMA per element)
//   gram_partial / chol_factor / apply block orthonormalisation by Cholesky QR (X^T X -> R^-1 -> X R^-1) and the
//                                      Rayleigh-Ritz products X^T Y; partial sums per row slab are combined in a fixed
//                                      order (deterministic)
//   rotate_residual_kernel             X <- X Q, Y <- Y Q and the residual norms |Y q_j - lambda_j X q_j| per Ritz pair
//   kmeans_lloyd_kernel                Lloyd iterations of sklearn's _kmeans_single_lloyd, one CTA per initialisation
//                                      (all n_init runs in one launch), fixed-order reductions
// The iteration control (power steps, Rayleigh-Ritz on the 8 x 8 projected matrix, convergence test, k-means++ seeding
// with sklearn's own routine and random stream, best-of-n_init) is host code in hvb/spectral.py.
#include "hvb_common.cuh"

#include <algorithm>
#include <math.h>

namespace {

constexpr int kB = 8;                 // vectors per block (>= n_clusters; the rest are guard vectors)
constexpr int kMvThreads = 256;       // 8 warps
constexpr int kMvRowsPerWarp = 4;
constexpr int kMvRows = (kMvThreads / 32) * kMvRowsPerWarp;     // 32 rows per CTA
constexpr int kMvTile = 512;          // columns of X staged per step: 8 x 512 doubles = 32 KB

// ---------------------------------------------------------------------------------------------- Laplacian pieces
// 32 columns per CTA, 8 warps: warp r adds rows r, r + 8, ... of its columns (a coalesced 256-byte segment per row, four
// loads in flight), the eight partial sums are combined in warp order — deterministic, not numpy's row-after-row order
// (the two differ in the last bits only; the solver's tolerance is 1e-10).
__global__ void __launch_bounds__(256)
degree_kernel(const double* __restrict__ a, int n, double* __restrict__ dd) {
    __shared__ double part[8][32];
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
    if (j < n) {
        int i = r;
        for (; i + 24 < n; i += 32) {
            const double v0 = a[(size_t)i * n + j], v1 = a[(size_t)(i + 8) * n + j], v2 = a[(size_t)(i + 16) * n + j], v3 = a[(size_t)(i + 24) * n + j];
            w0 = __dadd_rn(w0, i == j ? 0.0 : v0);
            w1 = __dadd_rn(w1, i + 8 == j ? 0.0 : v1);
            w2 = __dadd_rn(w2, i + 16 == j ? 0.0 : v2);
            w3 = __dadd_rn(w3, i + 24 == j ? 0.0 : v3);
        }
        for (; i < n; i += 8) w0 = __dadd_rn(w0, i == j ? 0.0 : a[(size_t)i * n + j]);
    }
    part[r][lane] = __dadd_rn(__dadd_rn(w0, w1), __dadd_rn(w2, w3));
    __syncthreads();
    if (r == 0 && j < n) {
        double w = 0.0;
        for (int k = 0; k < 8; k++) w = __dadd_rn(w, part[k][lane]);
        dd[j] = w == 0.0 ? 1.0 : sqrt(w);
    }
}

__global__ void __launch_bounds__(256)
normalize_kernel(const double* __restrict__ a, const double* __restrict__ dd, int n, double* __restrict__ m) {
    const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const size_t p = (size_t)i * n + j;
    m[p] = i == j ? 0.0 : __ddiv_rn(__ddiv_rn(a[p], dd[j]), dd[i]);
}

__global__ void __launch_bounds__(32)
column_sum_kernel(const double* __restrict__ partial, int n_blk, int width, double* __restrict__ out) {
    if (threadIdx.x < width) {
        double s = 0.0;
        for (int k = 0; k < n_blk; k++) s = __dadd_rn(s, partial[(size_t)k * width + threadIdx.x]);
        out[threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------- Y = (M + shift I) X
// x, y: [kB][n] (each vector contiguous).  A warp owns 4 consecutive rows of M; the CTA stages a 512-column tile of the
// 8 vectors in shared memory and every lane walks the tile two columns at a time: 4 coalesced 512-byte row segments of M
// (128-bit loads, 8 in flight per lane) against 8 shared-memory operand pairs = 64 DFMA per 4 loads.  blockIdx.y splits
// the columns so that small matrices still fill the 148 SMs; the split partial sums are combined in split order by
// matvec_finish_kernel (deterministic), which also adds shift * x.
template <bool VEC2>
__global__ void __launch_bounds__(kMvThreads)
sym_block_matvec_kernel(const double* __restrict__ m, int n, const double* __restrict__ x, int cols_per_split,
                        double* __restrict__ ypart) {
    __shared__ __align__(16) double xs[kB][kMvTile];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kMvRows + warp * kMvRowsPerWarp;
    const int c_begin = blockIdx.y * cols_per_split, c_end = min(n, c_begin + cols_per_split);
    double acc[kMvRowsPerWarp][kB];
#pragma unroll
    for (int r = 0; r < kMvRowsPerWarp; r++)
#pragma unroll
        for (int j = 0; j < kB; j++) acc[r][j] = 0.0;
    const double* mrow[kMvRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kMvRowsPerWarp; r++) mrow[r] = m + (size_t)min(row0 + r, n - 1) * n;

    for (int c0 = c_begin; c0 < c_end; c0 += kMvTile) {
        const int cw = min(kMvTile, c_end - c0);
        __syncthreads();
        for (int p = threadIdx.x; p < kB * kMvTile; p += kMvThreads) {
            const int j = p / kMvTile, c = p - j * kMvTile;
            xs[j][c] = c < cw ? x[(size_t)j * n + c0 + c] : 0.0;
        }
        __syncthreads();
        if (VEC2) {                                           // n even, c0 even: every (row, c0 + 2k) address is 16-byte aligned
            const int cw2 = cw & ~1;
#pragma unroll 1
            for (int c = 2 * lane; c < cw2; c += 64) {
                double2 mv[kMvRowsPerWarp];
#pragma unroll
                for (int r = 0; r < kMvRowsPerWarp; r++) mv[r] = __ldg(reinterpret_cast<const double2*>(mrow[r] + c0 + c));
#pragma unroll
                for (int j = 0; j < kB; j++) {
                    const double2 xv = *reinterpret_cast<const double2*>(&xs[j][c]);
#pragma unroll
                    for (int r = 0; r < kMvRowsPerWarp; r++) acc[r][j] = __fma_rn(mv[r].y, xv.y, __fma_rn(mv[r].x, xv.x, acc[r][j]));
                }
            }
            if ((cw & 1) && lane == 0) {                      // odd tail column of the last tile of a split
                const int c = cw - 1;
#pragma unroll
                for (int j = 0; j < kB; j++)
#pragma unroll
                    for (int r = 0; r < kMvRowsPerWarp; r++) acc[r][j] = __fma_rn(__ldg(mrow[r] + c0 + c), xs[j][c], acc[r][j]);
            }
        } else {
            for (int c = lane; c < cw; c += 32) {
                double mv[kMvRowsPerWarp];
#pragma unroll
                for (int r = 0; r < kMvRowsPerWarp; r++) mv[r] = __ldg(mrow[r] + c0 + c);
#pragma unroll
                for (int j = 0; j < kB; j++) {
                    const double xv = xs[j][c];
#pragma unroll
                    for (int r = 0; r < kMvRowsPerWarp; r++) acc[r][j] = __fma_rn(mv[r], xv, acc[r][j]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kMvRowsPerWarp; r++)
#pragma unroll
        for (int j = 0; j < kB; j++) {
            double v = acc[r][j];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
            acc[r][j] = v;
        }
    if (lane < kB) {
        double* yp = ypart + (size_t)blockIdx.y * kB * n;
#pragma unroll
        for (int r = 0; r < kMvRowsPerWarp; r++) {
            const int row = row0 + r;
            if (row < n) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < kB; j++) if (lane == j) v = acc[r][j];
                yp[(size_t)lane * n + row] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
matvec_finish_kernel(const double* __restrict__ ypart, int splits, int n, const double* __restrict__ x, double shift,
                     double* __restrict__ y) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kB * n) return;
    double s = 0.0;
    for (int k = 0; k < splits; k++) s = __dadd_rn(s, ypart[(size_t)k * kB * n + p]);
    y[p] = __fma_rn(shift, x[p], s);
}

// ---------------------------------------------------------------------------------------------- small products
// partial[blk][i * kB + j] = sum over the slab's rows of a[i][row] * b[j][row]; slab = 256 rows.
constexpr int kSlab = 256;

__global__ void __launch_bounds__(256)
gram_partial_kernel(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ partial) {
    __shared__ double as[kB][kSlab], bs[kB][kSlab];
    const int r0 = blockIdx.x * kSlab;
    for (int p = threadIdx.x; p < kB * kSlab; p += 256) {
        const int j = p / kSlab, r = p - j * kSlab;
        const bool ok = r0 + r < n;
        as[j][r] = ok ? a[(size_t)j * n + r0 + r] : 0.0;
        bs[j][r] = ok ? b[(size_t)j * n + r0 + r] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int pair = warp; pair < kB * kB; pair += 8) {
        const int i = pair / kB, j = pair - i * kB;
        double s = 0.0;
        for (int r = lane; r < kSlab; r += 32) s = __fma_rn(as[i][r], bs[j][r], s);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, off));
        if (lane == 0) partial[(size_t)blockIdx.x * kB * kB + pair] = s;
    }
}

// out[kB*kB] = sum of the partials in slab order.  mode 1: additionally Cholesky-factor the (symmetric) sum G = R^T R and
// overwrite out with R^-1 (upper triangular, row-major) for the Cholesky-QR step; out[kB*kB] = 1 if G was not positive
// definite (a collapsed block), else 0.
__global__ void __launch_bounds__(64)
small_reduce_kernel(const double* __restrict__ partial, int n_blk, int mode, double* __restrict__ out) {
    __shared__ double g[kB * kB];
    if (threadIdx.x < kB * kB) {
        double s = 0.0;
        for (int k = 0; k < n_blk; k++) s = __dadd_rn(s, partial[(size_t)k * kB * kB + threadIdx.x]);
        g[threadIdx.x] = s;
    }
    __syncthreads();
    if (mode == 0) {
        if (threadIdx.x < kB * kB) out[threadIdx.x] = g[threadIdx.x];
        return;
    }
    if (threadIdx.x == 0) {
        double r[kB * kB], inv[kB * kB];
        int bad = 0;
        for (int i = 0; i < kB * kB; i++) { r[i] = 0.0; inv[i] = 0.0; }
        for (int j = 0; j < kB; j++) {                       // R upper: G = R^T R
            double d = g[j * kB + j];
            for (int k = 0; k < j; k++) d -= r[k * kB + j] * r[k * kB + j];
            if (!(d > 0.0)) { bad = 1; d = 1.0; }
            d = sqrt(d);
            r[j * kB + j] = d;
            for (int i = j + 1; i < kB; i++) {
                double s = g[j * kB + i];
                for (int k = 0; k < j; k++) s -= r[k * kB + j] * r[k * kB + i];
                r[j * kB + i] = s / d;
            }
        }
        for (int j = 0; j < kB; j++) {                       // inv = R^-1 by back substitution, column by column
            inv[j * kB + j] = 1.0 / r[j * kB + j];
            for (int i = j - 1; i >= 0; i--) {
                double s = 0.0;
                for (int k = i + 1; k <= j; k++) s += r[i * kB + k] * inv[k * kB + j];
                inv[i * kB + j] = -s / r[i * kB + i];
            }
        }
        for (int i = 0; i < kB * kB; i++) out[i] = inv[i];
        out[kB * kB] = (double)bad;
    }
}

// x <- x Q (and y <- y Q when y != NULL) with q [kB][kB] row-major: new_j = sum_i old_i * q[i][j].  With lambda != NULL
// also res_partial[blk][j] = sum over the slab of (y_new_j - lambda_j x_new_j)^2.
__global__ void __launch_bounds__(256)
rotate_residual_kernel(double* __restrict__ x, double* __restrict__ y, int n, const double* __restrict__ q,
                       const double* __restrict__ lambda, double* __restrict__ res_partial) {
    __shared__ double qs[kB * kB], ls[kB], red[8][kB];
    if (threadIdx.x < kB * kB) qs[threadIdx.x] = q[threadIdx.x];
    if (threadIdx.x < kB) ls[threadIdx.x] = lambda ? lambda[threadIdx.x] : 0.0;
    __syncthreads();
    const int row = blockIdx.x * 256 + threadIdx.x;
    double rs[kB];
#pragma unroll
    for (int j = 0; j < kB; j++) rs[j] = 0.0;
    if (row < n) {
        double xo[kB], yo[kB];
#pragma unroll
        for (int i = 0; i < kB; i++) { xo[i] = x[(size_t)i * n + row]; yo[i] = y ? y[(size_t)i * n + row] : 0.0; }
#pragma unroll
        for (int j = 0; j < kB; j++) {
            double xn = 0.0, yn = 0.0;
#pragma unroll
            for (int i = 0; i < kB; i++) { xn = __fma_rn(xo[i], qs[i * kB + j], xn); yn = __fma_rn(yo[i], qs[i * kB + j], yn); }
            x[(size_t)j * n + row] = xn;
            if (y) y[(size_t)j * n + row] = yn;
            const double d = __fma_rn(-ls[j], xn, yn);
            rs[j] = d * d;
        }
    }
    if (!res_partial) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < kB; j++) {
        double v = rs[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < kB) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s = __dadd_rn(s, red[w][threadIdx.x]);
        res_partial[(size_t)blockIdx.x * kB + threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------- k-means (Lloyd)
// sklearn/cluster/_kmeans.py::_kmeans_single_lloyd on x [n][d] (already mean-centred by the caller), one CTA per
// initialisation.  E-step: label = argmin_c (|c|^2 - 2 x.c), first minimum wins; M-step: centre = mean of its points
// (an empty cluster sets flag 1 and stops: the host falls back to sklearn's relocation logic for that call); stop when
// the labels repeat (strict convergence) or sum |shift|^2 <= tol, else after max_iter; if not strictly converged the
// labels are recomputed against the final centres; inertia = sum |x - centre[label]|^2.
constexpr int kKmThreads = 256;
constexpr int kKmMaxK = 8, kKmMaxD = 8;

__device__ __forceinline__ double km_block_sum(double v, double* red /*[8]*/) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < kKmThreads / 32; w++) s = __dadd_rn(s, red[w]);
    return s;                                               // every thread returns the same fixed-order sum
}

__global__ void __launch_bounds__(kKmThreads)
kmeans_lloyd_kernel(const double* __restrict__ x, int n, int d, int k, const double* __restrict__ init_centers,
                    int max_iter, double tol, int32_t* __restrict__ labels_all, double* __restrict__ centers_out,
                    double* __restrict__ inertia_out, int32_t* __restrict__ n_iter_out, int32_t* __restrict__ flags_out) {
    __shared__ double cen[kKmMaxK * kKmMaxD], cnew[kKmMaxK * kKmMaxD], cn2[kKmMaxK], red[8];
    __shared__ int s_changed;
    const int run = blockIdx.x;
    int32_t* labels = labels_all + (size_t)run * n;
    for (int p = threadIdx.x; p < k * d; p += kKmThreads) cen[p] = init_centers[(size_t)run * k * d + p];
    for (int p = threadIdx.x; p < n; p += kKmThreads) labels[p] = -1;
    __syncthreads();

    auto assign = [&](bool count_changes) -> int {          // E-step; returns (block-wide) whether any label changed
        if (threadIdx.x < k) {
            double s = 0.0;
            for (int t = 0; t < d; t++) s = __fma_rn(cen[threadIdx.x * d + t], cen[threadIdx.x * d + t], s);
            cn2[threadIdx.x] = s;
        }
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        int changed = 0;
        for (int p = threadIdx.x; p < n; p += kKmThreads) {
            double best = 0.0;
            int bi = 0;
            for (int c = 0; c < k; c++) {
                double dot = 0.0;
                for (int t = 0; t < d; t++) dot = __fma_rn(x[(size_t)p * d + t], cen[c * d + t], dot);
                const double v = __fma_rn(-2.0, dot, cn2[c]);
                if (c == 0 || v < best) { best = v; bi = c; }
            }
            if (labels[p] != bi) changed = 1;
            labels[p] = bi;
        }
        if (count_changes && changed) s_changed = 1;
        __syncthreads();
        return s_changed;
    };

    int it = 0, strict = 0, flag = 0;
    for (; it < max_iter; it++) {
        const int changed = assign(true);
        // M-step from the labels just computed: fixed-order block sums per (cluster, dim)
        for (int c = 0; c < k; c++) {
            double cnt = 0.0, sums[kKmMaxD];
            for (int t = 0; t < kKmMaxD; t++) sums[t] = 0.0;
            for (int p = threadIdx.x; p < n; p += kKmThreads)
                if (labels[p] == c) {
                    cnt += 1.0;
                    for (int t = 0; t < d; t++) sums[t] = __dadd_rn(sums[t], x[(size_t)p * d + t]);
                }
            const double total = km_block_sum(cnt, red);
            for (int t = 0; t < d; t++) {
                const double s = km_block_sum(sums[t], red);
                if (threadIdx.x == 0) cnew[c * d + t] = total > 0.0 ? __ddiv_rn(s, total) : cen[c * d + t];
            }
            if (total == 0.0) flag = 1;
        }
        __syncthreads();
        double shift = 0.0;
        if (threadIdx.x == 0) {
            for (int p = 0; p < k * d; p++) { const double dl = cnew[p] - cen[p]; shift = __fma_rn(dl, dl, shift); }
            red[0] = shift;
        }
        __syncthreads();
        shift = red[0];
        __syncthreads();
        for (int p = threadIdx.x; p < k * d; p += kKmThreads) cen[p] = cnew[p];      // centers, centers_new = centers_new, centers
        __syncthreads();
        if (flag) { it++; break; }
        if (!changed) { strict = 1; it++; break; }
        if (shift <= tol) { it++; break; }
    }
    if (!strict && !flag) assign(false);
    // inertia against the final centres
    double part = 0.0;
    for (int p = threadIdx.x; p < n; p += kKmThreads) {
        const int c = labels[p];
        for (int t = 0; t < d; t++) { const double dl = x[(size_t)p * d + t] - cen[c * d + t]; part = __fma_rn(dl, dl, part); }
    }
    const double inertia = km_block_sum(part, red);
    if (threadIdx.x == 0) { inertia_out[run] = inertia; n_iter_out[run] = it; flags_out[run] = flag; }
    for (int p = threadIdx.x; p < k * d; p += kKmThreads) centers_out[(size_t)run * k * d + p] = cen[p];
}

}  // namespace

extern "C" {

int hvb_spectral_block(int* out_block) {
    HVB_ARG(out_block != nullptr, "null pointer");
    *out_block = kB;
    return HVB_OK;
}

int hvb_laplacian_normalize(hvb_ctx* ctx, const double* a_dev, int n, double* m_dev, double* dd_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(a_dev && m_dev && dd_dev && n >= 1 && n <= 65535, "bad arguments");
    degree_kernel<<<hvb_div_up(n, 32), 256, 0, ctx->stream>>>(a_dev, n, dd_dev);
    HVB_LAUNCHED(ctx);
    normalize_kernel<<<dim3(hvb_div_up(n, 256), n), 256, 0, ctx->stream>>>(a_dev, dd_dev, n, m_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_sym_block_matvec(hvb_ctx* ctx, const double* m_dev, int n, const double* x_dev, double shift, double* y_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(m_dev && x_dev && y_dev && n >= 1 && x_dev != y_dev, "bad arguments");
    const int row_ctas = hvb_div_up(n, kMvRows);
    int splits = std::max(1, std::min(8, hvb_div_up(2 * ctx->sm_count, row_ctas)));
    int cols = hvb_div_up(n, splits);
    cols = hvb_div_up(cols, 64) * 64;                         // split boundaries on 512-byte lines (and even columns)
    splits = hvb_div_up(n, cols);
    void* ypart = nullptr;
    HVB_TRY(hvb_scratch2(ctx, (size_t)splits * kB * n * sizeof(double), &ypart));
    const bool vec2 = (n % 2 == 0) && ((uintptr_t)m_dev % 16 == 0);
    const dim3 grid(row_ctas, splits);
    if (vec2) sym_block_matvec_kernel<true><<<grid, kMvThreads, 0, ctx->stream>>>(m_dev, n, x_dev, cols, (double*)ypart);
    else sym_block_matvec_kernel<false><<<grid, kMvThreads, 0, ctx->stream>>>(m_dev, n, x_dev, cols, (double*)ypart);
    HVB_LAUNCHED(ctx);
    matvec_finish_kernel<<<hvb_div_up((int64_t)kB * n, 256), 256, 0, ctx->stream>>>((const double*)ypart, splits, n, x_dev, shift, y_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

// out_dev[64] = A^T B for the two blocks a, b [8][n] (mode 0); mode 1 (a == b): out_dev[65] = R^-1 of the Cholesky
// factor of A^T A + the not-positive-definite flag.
int hvb_block_gram(hvb_ctx* ctx, const double* a_dev, const double* b_dev, int n, int mode, double* out_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(a_dev && b_dev && out_dev && n >= 1 && (mode == 0 || mode == 1), "bad arguments");
    const int n_blk = hvb_div_up(n, kSlab);
    void* partial = nullptr;
    HVB_TRY(hvb_scratch2(ctx, (size_t)n_blk * kB * kB * sizeof(double), &partial));
    gram_partial_kernel<<<n_blk, 256, 0, ctx->stream>>>(a_dev, b_dev, n, (double*)partial);
    HVB_LAUNCHED(ctx);
    small_reduce_kernel<<<1, 64, 0, ctx->stream>>>((const double*)partial, n_blk, mode, out_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

// x <- x Q (y <- y Q when y_dev != NULL); with lambda_dev: out_res_dev[8] = |y q_j - lambda_j x q_j|^2.
int hvb_block_rotate(hvb_ctx* ctx, double* x_dev, double* y_dev, int n, const double* q_dev, const double* lambda_dev,
                     double* out_res_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(x_dev && q_dev && n >= 1, "bad arguments");
    HVB_ARG((lambda_dev == nullptr) == (out_res_dev == nullptr), "lambda and out_res go together");
    HVB_ARG(lambda_dev == nullptr || y_dev != nullptr, "residuals need y");
    const int n_blk = hvb_div_up(n, 256);
    void* partial = nullptr;
    if (out_res_dev) HVB_TRY(hvb_scratch3(ctx, (size_t)n_blk * kB * kB * sizeof(double), &partial));
    rotate_residual_kernel<<<n_blk, 256, 0, ctx->stream>>>(x_dev, y_dev, n, q_dev, lambda_dev, (double*)partial);
    HVB_LAUNCHED(ctx);
    if (out_res_dev) {
        column_sum_kernel<<<1, 32, 0, ctx->stream>>>((const double*)partial, n_blk, kB, out_res_dev);
        HVB_LAUNCHED(ctx);
    }
    return HVB_OK;
}

int hvb_kmeans_lloyd(hvb_ctx* ctx, const double* x_dev, int n, int d, int k, const double* init_centers_dev, int n_init,
                     int max_iter, double tol, int32_t* out_labels_dev, double* out_centers_dev, double* out_inertia_dev,
                     int32_t* out_n_iter_dev, int32_t* out_flags_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(x_dev && init_centers_dev && out_labels_dev && out_centers_dev && out_inertia_dev && out_n_iter_dev && out_flags_dev, "null pointer");
    HVB_ARG(n >= 1 && n_init >= 1 && max_iter >= 1, "bad sizes");
    if (k < 1 || k > kKmMaxK || d < 1 || d > kKmMaxD) {
        hvb_set_error("hvb_kmeans_lloyd: k = %d, d = %d outside the kernel's range (<= %d clusters, <= %d dims)", k, d, kKmMaxK, kKmMaxD);
        return HVB_ERR_UNSUPPORTED;
    }
    kmeans_lloyd_kernel<<<n_init, kKmThreads, 0, ctx->stream>>>(x_dev, n, d, k, init_centers_dev, max_iter, tol, out_labels_dev,
                                                              out_centers_dev, out_inertia_dev, out_n_iter_dev, out_flags_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

}  // extern "C"

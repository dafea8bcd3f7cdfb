// K4a (float64 part) — StandardScaler.fit_transform (hockey/common/team_hybrid.py:166, :252) and the
// RBF affinity SpectralClustering(affinity='rbf', gamma=1.0) builds in fit (team_hybrid.py:185-193;
// sklearn pairwise: d2 = max(|x|^2+|y|^2-2x.y, 0), zero diagonal, A = exp(-gamma d2)).
//
// Two device paths produce d2 / A (hvb_gram_affinity `mode`):
//   mode 1  float64 CUDA cores only: d2_ij = sum_k (x_ik - x_jk)^2 accumulated directly (no
//           cancellation), 64x64 tiles, symmetric half computed and mirrored.
//   mode 0  tensor cores: G = X.X^T by the tcgen05 kernel in k4_gram_tcgen05.cu (split-TF32 operands,
//           fp32 accumulation in TMEM), d2 from G and float64 row norms, then every pair whose
//           affinity does not underflow (gamma*d2 below the refine cut) is recomputed here in
//           float64 — the tolerance on A (1e-3 relative) is an absolute 1e-3 on d2 ~ 1250, beyond
//           what an fp32-accumulated Gram guarantees for near-duplicate rows.
#include "hvb_common.cuh"

#include <math.h>

int hvb_gram_tc_launch(hvb_ctx* ctx, const double* x_dev, int n, int d, float* out_g_dev);   // k4_gram_tcgen05.cu

namespace {

// ------------------------------------------------------------------ StandardScaler
// grid.x = ceil(D/32); block (32, 8).  Mirrors sklearn's _incremental_mean_and_var for one batch:
// mean = sum/N; var = (sum((x-mean)^2) - (sum(x-mean))^2/N)/N; constant features -> scale 1.
__global__ void __launch_bounds__(256)
standardize_stats_kernel(const double* __restrict__ x, int n, int d, double* __restrict__ mean, double* __restrict__ scale) {
    __shared__ double s_a[8][33];
    __shared__ double s_b[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    double sum = 0.0;
    if (col < d) for (int r = threadIdx.y; r < n; r += 8) sum = __dadd_rn(sum, x[(int64_t)r * d + col]);
    s_a[threadIdx.y][threadIdx.x] = sum;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < 8; k++) tot = __dadd_rn(tot, s_a[k][threadIdx.x]);
    const double mu = __ddiv_rn(tot, (double)n);
    __syncthreads();
    double c1 = 0.0, c2 = 0.0;
    if (col < d) for (int r = threadIdx.y; r < n; r += 8) {
        double t = __dsub_rn(x[(int64_t)r * d + col], mu);
        c1 = __dadd_rn(c1, t);
        c2 = __dadd_rn(c2, __dmul_rn(t, t));
    }
    s_a[threadIdx.y][threadIdx.x] = c1;
    s_b[threadIdx.y][threadIdx.x] = c2;
    __syncthreads();
    if (threadIdx.y == 0 && col < d) {
        double corr = 0.0, ss = 0.0;
        for (int k = 0; k < 8; k++) { corr = __dadd_rn(corr, s_a[k][threadIdx.x]); ss = __dadd_rn(ss, s_b[k][threadIdx.x]); }
        const double dn = (double)n;
        double var = __ddiv_rn(__dsub_rn(ss, __ddiv_rn(__dmul_rn(corr, corr), dn)), dn);
        // sklearn _is_constant_feature: var <= n*eps*var + (n*mean*eps)^2
        const double eps = 2.220446049250313e-16;
        const double nme = __dmul_rn(__dmul_rn(dn, mu), eps);
        const double bound = __dadd_rn(__dmul_rn(__dmul_rn(dn, eps), var), __dmul_rn(nme, nme));
        double sc = sqrt(var);
        if (var <= bound) sc = 1.0;
        mean[col] = mu;
        scale[col] = sc;
    }
}

__global__ void __launch_bounds__(256)
scale_transform_kernel(const double* __restrict__ x, int64_t total, int d, const double* __restrict__ mean,
                       const double* __restrict__ scale, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % d);
        out[i] = __ddiv_rn(__dsub_rn(x[i], mean[c]), scale[c]);
    }
}

// ------------------------------------------------------------------ float64 squared distances
constexpr int kT = 64;      // output tile
constexpr int kK = 16;      // k-chunk

__global__ void __launch_bounds__(256)
d2_f64_kernel(const double* __restrict__ x, int n, int d, double gamma, double* __restrict__ out_d2, double* __restrict__ out_a) {
    // upper-triangular tile index -> (bi, bj), bj >= bi
    const int nb = (n + kT - 1) / kT;
    int t = blockIdx.x, bi = 0;
    while (t >= nb - bi) { t -= nb - bi; bi++; }
    const int bj = bi + t;
    __shared__ double sa[kK][kT + 1];
    __shared__ double sb[kK][kT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 4x4 outputs each
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < d; k0 += kK) {
        for (int e = threadIdx.x; e < kT * kK; e += 256) {
            const int r = e / kK, k = e % kK;
            const int gi = bi * kT + r, gj = bj * kT + r, gk = k0 + k;
            sa[k][r] = (gi < n && gk < d) ? x[(int64_t)gi * d + gk] : 0.0;
            sb[k][r] = (gj < n && gk < d) ? x[(int64_t)gj * d + gk] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kK; k++) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = sa[k][ty * 4 + i]; b[i] = sb[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) { double df = a[i] - b[j]; acc[i][j] = fma(df, df, acc[i][j]); }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gi = bi * kT + ty * 4 + i, gj = bj * kT + tx * 4 + j;
            if (gi < n && gj < n) {
                const double v = (gi == gj) ? 0.0 : acc[i][j];
                const double av = exp(-gamma * v);
                if (out_d2) { out_d2[(int64_t)gi * n + gj] = v; out_d2[(int64_t)gj * n + gi] = v; }
                if (out_a) { out_a[(int64_t)gi * n + gj] = av; out_a[(int64_t)gj * n + gi] = av; }
            }
        }
}

// ------------------------------------------------------------------ tensor-core Gram -> d2 / A + float64 refinement
__global__ void __launch_bounds__(256)
row_norms_kernel(const double* __restrict__ x, int n, int d, double* __restrict__ norms) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    double s = 0.0;
    for (int k = lane; k < d; k += 32) { double v = x[(int64_t)row * d + k]; s = fma(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norms[row] = s;
}

// One warp per output pair (i <= j): d2 from the fp32 Gram; pairs below the refine cut are
// recomputed exactly in float64 by the whole warp.
__global__ void __launch_bounds__(256)
affinity_from_gram_kernel(const double* __restrict__ x, const float* __restrict__ g, const double* __restrict__ norms, int n,
                          int d, double gamma, double refine_cut, double* __restrict__ out_d2, double* __restrict__ out_a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = (int64_t)n * n;
    // each warp takes 32 consecutive cells of the full matrix; lanes = cells; refinement is warp-cooperative
    for (int64_t c0 = warp_global * 32; c0 < total; c0 += n_warps * 32) {
        const int64_t c = c0 + lane;
        const bool valid = c < total;
        int i = 0, j = 0;
        double v = 0.0;
        bool need = false;
        if (valid) {
            i = (int)(c / n); j = (int)(c % n);
            const double gg = (double)g[c];
            v = fmax(norms[i] + norms[j] - 2.0 * gg, 0.0);
            if (i == j) v = 0.0;
            need = (i != j) && (gamma * v < refine_cut);
        }
        unsigned m = __ballot_sync(0xffffffffu, need);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int ri = __shfl_sync(0xffffffffu, i, src), rj = __shfl_sync(0xffffffffu, j, src);
            double s = 0.0;
            for (int k = lane; k < d; k += 32) {
                double df = x[(int64_t)ri * d + k] - x[(int64_t)rj * d + k];
                s = fma(df, df, s);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == src) v = s;
        }
        if (valid) {
            if (out_d2) out_d2[c] = v;
            if (out_a) out_a[c] = exp(-gamma * v);
        }
    }
}

}  // namespace

extern "C" {

int hvb_standardize(hvb_ctx* ctx, const double* x_dev, int n, int d, double* out_mean_dev, double* out_scale_dev,
                    double* out_xs_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n > 0 && d > 0, "empty matrix");
    HVB_ARG(x_dev && out_mean_dev && out_scale_dev, "null pointer");
    standardize_stats_kernel<<<hvb_div_up(d, 32), dim3(32, 8), 0, ctx->stream>>>(x_dev, n, d, out_mean_dev, out_scale_dev);
    HVB_LAUNCHED(ctx);
    if (out_xs_dev) HVB_TRY(hvb_scale_transform(ctx, x_dev, n, d, out_mean_dev, out_scale_dev, out_xs_dev));
    return HVB_OK;
}

int hvb_scale_transform(hvb_ctx* ctx, const double* x_dev, int n, int d, const double* mean_dev, const double* scale_dev,
                        double* out_xs_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0 && d > 0, "bad shape");
    if (n == 0) return HVB_OK;
    HVB_ARG(x_dev && mean_dev && scale_dev && out_xs_dev, "null pointer");
    const int64_t total = (int64_t)n * d;
    int grid = (int)((total + 255) / 256);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    scale_transform_kernel<<<grid, 256, 0, ctx->stream>>>(x_dev, total, d, mean_dev, scale_dev, out_xs_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_gram_affinity(hvb_ctx* ctx, const double* x_dev, int n, int d, double gamma, int mode, double* out_d2_dev,
                      double* out_a_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0 && d > 0, "bad shape");
    HVB_ARG(mode == 0 || mode == 1, "mode must be 0 (tensor cores + refine) or 1 (float64)");
    if (n == 0 || (!out_d2_dev && !out_a_dev)) return HVB_OK;
    HVB_ARG(x_dev != nullptr, "null input");
    if (mode == 1) {
        const int nb = hvb_div_up(n, kT);
        d2_f64_kernel<<<nb * (nb + 1) / 2, 256, 0, ctx->stream>>>(x_dev, n, d, gamma, out_d2_dev, out_a_dev);
        HVB_LAUNCHED(ctx);
        return HVB_OK;
    }
    // mode 0: tcgen05 Gram + float64 epilogue
    uint8_t* s = nullptr;
    const size_t g_bytes = (((size_t)n * n * sizeof(float)) + 255) & ~(size_t)255;
    HVB_TRY(hvb_scratch2(ctx, g_bytes + (size_t)n * sizeof(double), (void**)&s));
    float* g = (float*)s;
    double* norms = (double*)(s + g_bytes);
    HVB_TRY(hvb_gram_tc_launch(ctx, x_dev, n, d, g));
    row_norms_kernel<<<hvb_div_up(n, 8), 256, 0, ctx->stream>>>(x_dev, n, d, norms);
    HVB_LAUNCHED(ctx);
    // exp(-t) underflows to 0 in float64 beyond t ~ 745; anything that could still be > 1e-300
    // (t < 690.8) sits far below this cut even after the fp32 Gram error.
    const double refine_cut = 760.0;
    int grid = (int)(((int64_t)n * n + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    affinity_from_gram_kernel<<<grid, 256, 0, ctx->stream>>>(x_dev, g, norms, n, d, gamma, refine_cut, out_d2_dev, out_a_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_gram_tc(hvb_ctx* ctx, const double* x_dev, int n, int d, float* out_g_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n > 0 && d > 0, "bad shape");
    HVB_ARG(x_dev && out_g_dev, "null pointer");
    return hvb_gram_tc_launch(ctx, x_dev, n, d, out_g_dev);
}

int hvb_gram_affinity_host(hvb_ctx* ctx, const double* x_host, int n, int d, double gamma, int mode, double* out_d2_host,
                           double* out_a_host) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0 && d > 0, "bad shape");
    if (n == 0) return HVB_OK;
    HVB_ARG(x_host != nullptr, "null input");
    const size_t xb = (((size_t)n * d * 8) + 255) & ~(size_t)255, mb = (size_t)n * n * 8;
    uint8_t* dv = nullptr;
    HVB_TRY(hvb_scratch(ctx, xb + 2 * mb, (void**)&dv));
    HVB_CUDA(cudaMemcpyAsync(dv, x_host, (size_t)n * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    double* d2 = out_d2_host ? (double*)(dv + xb) : nullptr;
    double* a = out_a_host ? (double*)(dv + xb + mb) : nullptr;
    HVB_TRY(hvb_gram_affinity(ctx, (const double*)dv, n, d, gamma, mode, d2, a));
    if (d2) HVB_CUDA(cudaMemcpyAsync(out_d2_host, d2, mb, cudaMemcpyDeviceToHost, ctx->stream));
    if (a) HVB_CUDA(cudaMemcpyAsync(out_a_host, a, mb, cudaMemcpyDeviceToHost, ctx->stream));
    HVB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HVB_OK;
}

}  // extern "C"

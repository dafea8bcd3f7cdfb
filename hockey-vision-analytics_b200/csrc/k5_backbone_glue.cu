// K5 — the element-wise glue between the library convolutions of the YOLOv8 / MobileNetV3 forwards
// (reached from hockey/main.py:179-184 and common/team_hybrid.py:73-81).  PyTorch runs the
// convolutions; everything between two convolutions is one pass over HBM here instead of the 3-6
// passes eager PyTorch makes (broadcast bias add, SiLU, residual add, chunk copy, cat, upsample):
//
//   bias_act_kernel       y = act(conv_raw + bias) (+ residual), written to up to two destinations
//                         (a contiguous NHWC tensor for the next convolution and/or a channel slice
//                         of a concat buffer) — ultralytics Conv.forward_fuse, Bottleneck.forward,
//                         and the torch.cat of C2f.forward / SPPF.forward / Detect.forward
//   concat_nhwc_kernel    channel concat of up to four NHWC sources, each optionally nearest-
//                         upsampled by 2^s — nn.Upsample(2,'nearest') + Concat of the yolov8.yaml neck
//   stem_conv_kernel      layer 0: 3x3 stride-2 convolution of the NCHW letterbox output (K1) straight
//                         to NHWC with bias + SiLU — exact fp32 FMA, weights as kernel parameters
//                         (constant bank operands), one block = 128 output pixels x 8 rows, the next
//                         row's input prefetched into registers while the current row's FMAs run
//   sppf_pool_concat_kernel  SPPF's three cascaded 5x5 max-pools + concat on an on-chip tile
//
// All three are HBM-bound streaming kernels: 128-bit accesses, 4 independent vectors in flight per
// thread, grid sized to the data.  Layout everywhere: NHWC ("channels_last") float32.
#include "hvb_common.cuh"

namespace {

enum { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2, ACT_HSWISH = 3, ACT_SILU_FAST = 4 };

// SiLU with ex2.approx + rcp.approx (about 2^-22 relative error each, <= 1e-6 relative in total where the output is not negligible — three orders of
// magnitude below the TF32 rounding of the convolution that follows) instead of expf + IEEE division: 6 instructions
// instead of ~35 per value.  The stem kernel is instruction-issue bound, and so is the in-place bias + SiLU epilogue
// with the exact formula (measured: 530 us -> 416 us on the 1.45 GB layer-0 tensor, 5.5 -> 7.0 TB/s).
__device__ __forceinline__ float silu_fast(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(v, -1.4426950408889634f)));
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(1.0f, e)));
    return __fmul_rn(v, r);
}

template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
    if (ACT == ACT_SILU_FAST) return silu_fast(v);
    if (ACT == ACT_SILU) return __fdiv_rn(v, __fadd_rn(1.0f, expf(-v)));          // x / (1 + exp(-x)), as torch's silu kernel
    if (ACT == ACT_RELU) return fmaxf(v, 0.0f);
    if (ACT == ACT_HSWISH) return __fdiv_rn(__fmul_rn(v, fminf(fmaxf(__fadd_rn(v, 3.0f), 0.0f), 6.0f)), 6.0f);
    return v;
}

template <int V> struct Vec;
template <> struct Vec<4> { typedef float4 T; };
template <> struct Vec<2> { typedef float2 T; };
template <> struct Vec<1> { typedef float T; };
struct __align__(32) float8 { float4 lo, hi; };
template <> struct Vec<8> { typedef float8 T; };

template <int V> __device__ __forceinline__ void vload(const float* p, float (&r)[V]) {
    typename Vec<V>::T t = *reinterpret_cast<const typename Vec<V>::T*>(p);
    const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
    for (int i = 0; i < V; ++i) r[i] = f[i];
}
template <int V> __device__ __forceinline__ void vstore(float* p, const float (&r)[V]) {
    typename Vec<V>::T t;
    float* f = reinterpret_cast<float*>(&t);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = r[i];
    *reinterpret_cast<typename Vec<V>::T*>(p) = t;
}

struct EpiArgs {
    const float* x;        // [npix, C] raw convolution output (NHWC, dense)
    const float* bias;     // [C] or null
    const float* res;      // [npix, C] dense residual or null
    float* out1;           // out1[p*ld1 + off1 + c], all C channels (may alias x), or null
    float* out2;           // out2[p*ld2 + off2 + (c - c2_begin)] for c in [c2_begin, c2_begin + c2_count), or null
    int64_t ld1, off1, ld2, off2;
    uint32_t total;        // npix * C / V
    uint32_t cv;           // C / V
    int c2_begin, c2_count;
    int up2_h, up2_w;      // > 0: out2 is a [n, 2h, 2w, ld2] tensor and every source pixel is replicated 2x2 (nn.Upsample nearest)
};

constexpr int EPI_THREADS = 256;
constexpr int EPI_UNROLL = 4;

template <int V, int ACT>
__global__ void __launch_bounds__(EPI_THREADS)
bias_act_kernel(const EpiArgs a) {
    const uint32_t base = blockIdx.x * (EPI_THREADS * EPI_UNROLL) + threadIdx.x;
    float v[EPI_UNROLL][V], r[EPI_UNROLL][V];
    uint32_t idx[EPI_UNROLL];
#pragma unroll
    for (int u = 0; u < EPI_UNROLL; ++u) {
        idx[u] = base + u * EPI_THREADS;
        if (idx[u] < a.total) vload<V>(a.x + (size_t)idx[u] * V, v[u]);
    }
    if (a.res) {
#pragma unroll
        for (int u = 0; u < EPI_UNROLL; ++u)
            if (idx[u] < a.total) vload<V>(a.res + (size_t)idx[u] * V, r[u]);
    }
#pragma unroll
    for (int u = 0; u < EPI_UNROLL; ++u) {
        if (idx[u] >= a.total) continue;
        const uint32_t p = idx[u] / a.cv;
        const int c = (int)(idx[u] - p * a.cv) * V;
        float b[V];
        if (a.bias) vload<V>(a.bias + c, b);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float t = a.bias ? __fadd_rn(v[u][i], b[i]) : v[u][i];
            t = act_fn<ACT>(t);
            if (a.res) t = __fadd_rn(r[u][i], t);                                   // Bottleneck: x + cv2(cv1(x))
            v[u][i] = t;
        }
        if (a.out1) vstore<V>(a.out1 + (size_t)p * a.ld1 + a.off1 + c, v[u]);
        if (a.out2) {
            const int c2 = c - a.c2_begin;
            if (c2 >= 0 && c2 < a.c2_count) {
                if (a.up2_w > 0) {
                    const uint32_t t = p / (uint32_t)a.up2_w, x = p - t * (uint32_t)a.up2_w;
                    const uint32_t n = t / (uint32_t)a.up2_h, y = t - n * (uint32_t)a.up2_h;
                    const size_t w2 = 2 * (size_t)a.up2_w;
                    float* q = a.out2 + (((size_t)n * 2 * a.up2_h + 2 * y) * w2 + 2 * x) * a.ld2 + a.off2 + c2;
                    vstore<V>(q, v[u]);
                    vstore<V>(q + a.ld2, v[u]);
                    vstore<V>(q + w2 * a.ld2, v[u]);
                    vstore<V>(q + (w2 + 1) * a.ld2, v[u]);
                } else {
                    vstore<V>(a.out2 + (size_t)p * a.ld2 + a.off2 + c2, v[u]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
struct CatArgs {
    const float* src[4];
    int c[4];              // channels of each source
    int shift[4];          // source spatial size = (H >> shift, W >> shift)
    int nsrc;
    int H, W, ctot;        // output [n, H, W, ctot]
    float* out;
};

template <int V>
__global__ void __launch_bounds__(256)
concat_nhwc_kernel(const CatArgs a) {
    const int row = blockIdx.x;                       // n * H + y
    const int n = row / a.H, y = row - n * a.H;
    const uint32_t cv = a.ctot / V;
    const uint32_t rowlen = (uint32_t)a.W * cv;
    float* orow = a.out + (size_t)row * a.W * a.ctot;
    for (uint32_t j0 = blockIdx.y * 1024 + threadIdx.x; j0 < rowlen; j0 += gridDim.y * 1024) {
        float t[4][V];
        size_t dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                  // four independent loads in flight per thread
            const uint32_t j = j0 + u * 256;
            if (j >= rowlen) continue;
            const uint32_t x = j / cv;
            int c = (int)(j - x * cv) * V;
            dst[u] = (size_t)x * a.ctot + c;
            int s = 0;
            while (s + 1 < a.nsrc && c >= a.c[s]) { c -= a.c[s]; ++s; }
            const int sh = a.shift[s];
            const int hs = a.H >> sh, ws = a.W >> sh;
            vload<V>(a.src[s] + (((size_t)n * hs + (y >> sh)) * ws + (x >> sh)) * a.c[s] + c, t[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (j0 + u * 256 < rowlen) vstore<V>(orow + dst[u], t[u]);
    }
}

// ------------------------------------------------------------------------------------------------
// Stem: out[n, oy, ox, co] = silu(bias[co] + sum_{ci,ky,kx} w[co][ci][ky][kx] * in[n, ci, 2oy-1+ky, 2ox-1+kx])
// Input NCHW (what K1 writes), zero padding 1.  Weights arrive as kernel parameters, i.e. in the
// constant bank: with the loops fully unrolled every FFMA takes its weight as a c[0][..] operand, so
// the inner loop is 27 shared-memory loads + 27*CO FFMAs per pixel and no weight traffic at all.
#ifndef HVB_STEM_PX
#define HVB_STEM_PX 128
#endif
constexpr int STEM_PX = HVB_STEM_PX;         // output pixels (one row segment) per block = threads per block
template <int CO> struct StemParams { float w[27 * CO]; float b[CO]; };   // w index: ((ci*3+ky)*3+kx)*CO + co

#ifndef HVB_STEM_ROWS
#define HVB_STEM_ROWS 8
#endif
constexpr int STEM_ROWS = HVB_STEM_ROWS;                 // output rows per block: amortises the block start-up and lets row r+1's loads fly under row r's FMAs

template <int CO>
__global__ void __launch_bounds__(STEM_PX)
stem_conv_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int OH, int OW,
                 const __grid_constant__ StemParams<CO> prm) {
    // Input window of one output row, de-interleaved by column parity so that the three taps of thread t are
    // stride-1 across the warp: window column j (frame column 2*ox0 - 2 + j) lives in s_ev[j/2] or s_od[j/2];
    // tap kx of output t reads window column 2t + kx + 1  ->  od[t], ev[t+1], od[t+1].
    __shared__ float s_ev[3][3][STEM_PX + 2];
    __shared__ float s_od[3][3][STEM_PX + 2];
    constexpr int LDO = CO + 4;                                   // 16-byte aligned rows, conflict-free 128-bit accesses
    __shared__ __align__(16) float s_out[STEM_PX * LDO];
    const int n = blockIdx.z, oy0 = blockIdx.y * STEM_ROWS, ox0 = blockIdx.x * STEM_PX;
    const int t = threadIdx.x;
    const int ixw = 2 * ox0 - 2;                                  // frame column of window column 0 (even)
    const bool pair_ok = (W & 1) == 0 && ((reinterpret_cast<uintptr_t>(in) & 7) == 0);
    const size_t hw = (size_t)H * W;
    const float* img = in + (size_t)n * 3 * hw;
    // Thread t owns window pair t (columns 2t, 2t+1) of the nine (row, channel) planes; threads 0 and 1 also own pairs
    // STEM_PX and STEM_PX+1.  Column validity is per thread, row validity per output row.
    const int ixa = ixw + 2 * t, ixb = ixw + 2 * (STEM_PX + t);
    const bool a0 = ixa >= 0 && ixa < W, a1 = ixa + 1 >= 0 && ixa + 1 < W;
    const bool b0 = t < 2 && ixb >= 0 && ixb < W, b1 = t < 2 && ixb + 1 >= 0 && ixb + 1 < W;

    auto fetch = [&](int oy, int ix, bool c0, bool c1, float2 (&v)[9]) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy - 1 + ky;
            const bool row_ok = iy >= 0 && iy < H;
            const float* p = img + (size_t)(row_ok ? iy : 0) * W + ix;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                float2 r = make_float2(0.0f, 0.0f);
                if (row_ok) {
                    if (pair_ok && c0 && c1) {
                        r = __ldg(reinterpret_cast<const float2*>(p + ci * hw));
                    } else {
                        if (c0) r.x = __ldg(p + ci * hw);
                        if (c1) r.y = __ldg(p + ci * hw + 1);
                    }
                }
                v[ky * 3 + ci] = r;
            }
        }
    };
    auto put = [&](int j, const float2 (&v)[9]) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) { s_ev[ci][ky][j] = v[ky * 3 + ci].x; s_od[ci][ky][j] = v[ky * 3 + ci].y; }
    };

    float2 cur[9], ext[9];
    fetch(oy0, ixa, a0, a1, cur);
    if (t < 2) fetch(oy0, ixb, b0, b1, ext);
    const int rows = min(STEM_ROWS, OH - oy0);
    for (int r = 0; r < rows; ++r) {
        const int oy = oy0 + r;
        put(t, cur);
        if (t < 2) put(STEM_PX + t, ext);
        if (r + 1 < rows) {                                       // next row's loads are in flight during this row's FMAs
            fetch(oy + 1, ixa, a0, a1, cur);
            if (t < 2) fetch(oy + 1, ixb, b0, b1, ext);
        }
        __syncthreads();
        float acc[CO];
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[co] = prm.b[co];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float x0 = s_od[ci][ky][t], x1 = s_ev[ci][ky][t + 1], x2 = s_od[ci][ky][t + 1];
#pragma unroll
                for (int co = 0; co < CO; ++co) {
                    acc[co] = __fmaf_rn(prm.w[((ci * 3 + ky) * 3 + 0) * CO + co], x0, acc[co]);
                    acc[co] = __fmaf_rn(prm.w[((ci * 3 + ky) * 3 + 1) * CO + co], x1, acc[co]);
                    acc[co] = __fmaf_rn(prm.w[((ci * 3 + ky) * 3 + 2) * CO + co], x2, acc[co]);
                }
            }
#pragma unroll
        for (int c4 = 0; c4 < CO / 4; ++c4) {
            float4 v;
            v.x = silu_fast(acc[4 * c4]); v.y = silu_fast(acc[4 * c4 + 1]); v.z = silu_fast(acc[4 * c4 + 2]); v.w = silu_fast(acc[4 * c4 + 3]);
            *reinterpret_cast<float4*>(&s_out[t * LDO + 4 * c4]) = v;
        }
        __syncthreads();                                          // s_out complete; every thread is done reading s_ev / s_od
        // the row segment's 128 x CO outputs are one contiguous NHWC run: coalesced 128-bit stores
        const int npx = min(STEM_PX, OW - ox0);
        float4* o4 = reinterpret_cast<float4*>(out + (((size_t)n * OH + oy) * OW + ox0) * CO);
        constexpr int Q = CO / 4;
#pragma unroll 4
        for (int i = t; i < npx * Q; i += STEM_PX) {
            const int px = i / Q, q = i - px * Q;
            o4[i] = *reinterpret_cast<const float4*>(&s_out[px * LDO + 4 * q]);
        }
        // next iteration: put() only touches s_ev / s_od (free since the barrier above); s_out is rewritten after the
        // next iteration's first barrier, which every thread reaches only after finishing the stores above
    }
}

// ------------------------------------------------------------------------------------------------
// SPPF pooling: ultralytics SPPF.forward computes y1 = m(y0), y2 = m(y1), y3 = m(y2) with m = MaxPool2d(5, 1, 2)
// and concatenates [y0, y1, y2, y3].  One CTA holds an (image, CH-channel slice) tile of y0 in shared memory and
// runs the three cascaded pools on chip (each separable: 5-wide row max, then 5-tall column max; out-of-range
// taps are skipped, which is MaxPool's implicit -inf padding), writing every stage straight into its slice of
// the NHWC concat buffer: y0 is read from HBM once and nothing but the concat buffer is written.
template <int CH>
__global__ void __launch_bounds__(256)
sppf_pool_concat_kernel(const float* __restrict__ y0, int h, int w, int c, float* __restrict__ cat) {
    extern __shared__ __align__(16) float sp[];                 // [2][h*w][CH]: current stage, row-max scratch
    typedef typename Vec<CH>::T VT;
    const int n = blockIdx.y, c0 = blockIdx.x * CH, hw = h * w;
    VT* cur = reinterpret_cast<VT*>(sp);
    VT* tmp = cur + hw;
    const float* src = y0 + (size_t)n * hw * c + c0;
    float* dst = cat + (size_t)n * hw * 4 * c + c0;
    for (int p = threadIdx.x; p < hw; p += 256) {
        const VT v = *reinterpret_cast<const VT*>(src + (size_t)p * c);
        cur[p] = v;
        *reinterpret_cast<VT*>(dst + (size_t)p * 4 * c) = v;     // slice 0: y0 itself
    }
    __syncthreads();
    for (int stage = 1; stage <= 3; ++stage) {
        for (int p = threadIdx.x; p < hw; p += 256) {             // row max over x-2..x+2
            const int y = p / w, x = p - y * w;
            float m[CH];
            const float* f = reinterpret_cast<const float*>(&cur[p]);
#pragma unroll
            for (int k = 0; k < CH; ++k) m[k] = f[k];
#pragma unroll
            for (int d = -2; d <= 2; ++d) {
                if (d == 0 || x + d < 0 || x + d >= w) continue;
                const float* g = reinterpret_cast<const float*>(&cur[p + d]);
#pragma unroll
                for (int k = 0; k < CH; ++k) m[k] = fmaxf(m[k], g[k]);
            }
            float* o = reinterpret_cast<float*>(&tmp[p]);
#pragma unroll
            for (int k = 0; k < CH; ++k) o[k] = m[k];
        }
        __syncthreads();
        for (int p = threadIdx.x; p < hw; p += 256) {             // column max over y-2..y+2, write stage slice
            const int y = p / w;
            float m[CH];
            const float* f = reinterpret_cast<const float*>(&tmp[p]);
#pragma unroll
            for (int k = 0; k < CH; ++k) m[k] = f[k];
#pragma unroll
            for (int d = -2; d <= 2; ++d) {
                if (d == 0 || y + d < 0 || y + d >= h) continue;
                const float* g = reinterpret_cast<const float*>(&tmp[p + d * w]);
#pragma unroll
                for (int k = 0; k < CH; ++k) m[k] = fmaxf(m[k], g[k]);
            }
            VT v;
            float* o = reinterpret_cast<float*>(&v);
#pragma unroll
            for (int k = 0; k < CH; ++k) o[k] = m[k];
            cur[p] = v;                                           // safe: this pass reads tmp only
            *reinterpret_cast<VT*>(dst + (size_t)p * 4 * c + (size_t)stage * c) = v;
        }
        __syncthreads();
    }
}

template <int V>
int launch_bias_act(hvb_ctx* ctx, const EpiArgs& a, int act) {
    const uint32_t per_block = EPI_THREADS * EPI_UNROLL;
    const uint32_t grid = (a.total + per_block - 1) / per_block;
    switch (act) {
        case ACT_NONE: bias_act_kernel<V, ACT_NONE><<<grid, EPI_THREADS, 0, ctx->stream>>>(a); break;
        case ACT_SILU: bias_act_kernel<V, ACT_SILU><<<grid, EPI_THREADS, 0, ctx->stream>>>(a); break;
        case ACT_RELU: bias_act_kernel<V, ACT_RELU><<<grid, EPI_THREADS, 0, ctx->stream>>>(a); break;
        case ACT_HSWISH: bias_act_kernel<V, ACT_HSWISH><<<grid, EPI_THREADS, 0, ctx->stream>>>(a); break;
        case ACT_SILU_FAST: bias_act_kernel<V, ACT_SILU_FAST><<<grid, EPI_THREADS, 0, ctx->stream>>>(a); break;
        default: hvb_set_error("hvb_bias_act: unknown activation %d", act); return HVB_ERR_ARG;
    }
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

template <int CO>
int launch_stem(hvb_ctx* ctx, const float* in, const float* w_host, const float* b_host, int n, int h, int w, float* out) {
    StemParams<CO> prm;
    // PyTorch weight layout [CO][3][3][3] (co, ci, ky, kx) -> tap-major so one tap's CO weights are adjacent
    for (int co = 0; co < CO; ++co)
        for (int k = 0; k < 27; ++k) prm.w[k * CO + co] = w_host[co * 27 + k];
    for (int co = 0; co < CO; ++co) prm.b[co] = b_host ? b_host[co] : 0.0f;
    const int oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;           // floor((h + 2 - 3) / 2) + 1
    dim3 grid((ow + STEM_PX - 1) / STEM_PX, (oh + STEM_ROWS - 1) / STEM_ROWS, n);
    stem_conv_kernel<CO><<<grid, STEM_PX, 0, ctx->stream>>>(in, out, h, w, oh, ow, prm);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

inline bool aligned_to(const void* p, int bytes) { return ((uintptr_t)p % bytes) == 0; }

}  // namespace

extern "C" {

int hvb_bias_act(hvb_ctx* ctx, const float* x_dev, const float* bias_dev, const float* residual_dev, int64_t npix,
                 int channels, int act, float* out1_dev, int64_t out1_ld, int64_t out1_off, float* out2_dev,
                 int64_t out2_ld, int64_t out2_off, int c2_begin, int c2_count, int out2_up2_h, int out2_up2_w) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(npix >= 0 && channels > 0, "bad sizes");
    if (npix == 0) return HVB_OK;
    HVB_ARG(x_dev && (out1_dev || out2_dev), "null pointer");
    HVB_ARG(!out1_dev || (out1_ld >= channels && out1_off >= 0 && out1_off + channels <= out1_ld), "out1 slice outside its row");
    HVB_ARG(!out2_dev || (c2_begin >= 0 && c2_count > 0 && c2_begin + c2_count <= channels && out2_off >= 0 &&
                          out2_off + c2_count <= out2_ld), "out2 slice outside its row");
    HVB_ARG(npix * (int64_t)channels < ((int64_t)1 << 32) - 4096, "tensor too large for one launch");
    EpiArgs a;
    a.x = x_dev; a.bias = bias_dev; a.res = residual_dev; a.out1 = out1_dev; a.out2 = out2_dev;
    a.ld1 = out1_ld; a.off1 = out1_off; a.ld2 = out2_ld; a.off2 = out2_off;
    a.c2_begin = c2_begin; a.c2_count = out2_dev ? c2_count : 0;
    a.up2_h = a.up2_w = 0;
    if (out2_dev && (out2_up2_h > 0 || out2_up2_w > 0)) {
        HVB_ARG(out2_up2_h > 0 && out2_up2_w > 0 && npix % ((int64_t)out2_up2_h * out2_up2_w) == 0,
                "upsampled destination: npix is not a multiple of h*w");
        a.up2_h = out2_up2_h; a.up2_w = out2_up2_w;
    }
    // widest vector every address involved is aligned to
    auto ok = [&](int v) {
        if (channels % v) return false;
        if (!aligned_to(x_dev, 4 * v) || (bias_dev && !aligned_to(bias_dev, 4 * v)) || (residual_dev && !aligned_to(residual_dev, 4 * v))) return false;
        if (out1_dev && (!aligned_to(out1_dev, 4 * v) || out1_ld % v || out1_off % v)) return false;
        if (out2_dev && (!aligned_to(out2_dev, 4 * v) || out2_ld % v || out2_off % v || c2_begin % v || c2_count % v)) return false;
        return true;
    };
    const int v = ok(4) ? 4 : ok(2) ? 2 : 1;
    a.cv = (uint32_t)(channels / v);
    a.total = (uint32_t)(npix * channels / v);
    if (v == 4) return launch_bias_act<4>(ctx, a, act);
    if (v == 2) return launch_bias_act<2>(ctx, a, act);
    return launch_bias_act<1>(ctx, a, act);
}

int hvb_concat_nhwc(hvb_ctx* ctx, const float* const src_dev[4], const int32_t src_channels[4], const int32_t src_shift[4],
                    int n_src, int n, int h, int w, float* out_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n_src >= 1 && n_src <= 4 && n >= 0 && h > 0 && w > 0 && out_dev, "bad arguments");
    if (n == 0) return HVB_OK;
    CatArgs a;
    a.nsrc = n_src; a.H = h; a.W = w; a.out = out_dev; a.ctot = 0;
    int v = 4;
    for (int s = 0; s < 4; ++s) { a.src[s] = nullptr; a.c[s] = 0; a.shift[s] = 0; }
    for (int s = 0; s < n_src; ++s) {
        HVB_ARG(src_dev[s] && src_channels[s] > 0 && src_shift[s] >= 0 && src_shift[s] < 8, "bad source");
        HVB_ARG((h % (1 << src_shift[s])) == 0 && (w % (1 << src_shift[s])) == 0, "output size not a multiple of the upsample factor");
        a.src[s] = src_dev[s]; a.c[s] = src_channels[s]; a.shift[s] = src_shift[s];
        a.ctot += src_channels[s];
        while (v > 1 && (src_channels[s] % v || !aligned_to(src_dev[s], 4 * v))) v >>= 1;
    }
    while (v > 1 && !aligned_to(out_dev, 4 * v)) v >>= 1;
    const int64_t rows = (int64_t)n * h;
    HVB_ARG(rows < ((int64_t)1 << 31), "n*h exceeds the grid's x extent; split the batch");
    const uint32_t rowlen = (uint32_t)w * (a.ctot / v);
    int gy = (int)((rowlen + 1023) / 1024);
    if (gy > 64) gy = 64;
    dim3 grid((unsigned)rows, gy);
    if (v == 4) concat_nhwc_kernel<4><<<grid, 256, 0, ctx->stream>>>(a);
    else if (v == 2) concat_nhwc_kernel<2><<<grid, 256, 0, ctx->stream>>>(a);
    else concat_nhwc_kernel<1><<<grid, 256, 0, ctx->stream>>>(a);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_sppf_pool_concat(hvb_ctx* ctx, const float* y0_dev, int n, int h, int w, int channels, float* out_cat_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(y0_dev && out_cat_dev && n >= 0 && h > 0 && w > 0 && channels > 0, "bad arguments");
    if (n == 0) return HVB_OK;
    HVB_ARG(n <= 65535, "more than 65535 images");
    HVB_ARG(channels % 4 == 0 && aligned_to(y0_dev, 16) && aligned_to(out_cat_dev, 16), "channels must be a multiple of 4 and buffers 16-byte aligned");
    const size_t per_ch = (size_t)2 * h * w * sizeof(float);
    if (per_ch * 4 > 200 * 1024) {
        hvb_set_error("hvb_sppf_pool_concat: a %dx%d map does not fit shared memory; use MaxPool2d + hvb_concat_nhwc", h, w);
        return HVB_ERR_UNSUPPORTED;
    }
    // 8-channel slices (32-byte sectors per pixel) when two tiles of them fit, 4-channel slices otherwise
    // (ctx->max_smem_optin is not assumed: 200 KB is within every sm_100 part's 227 KB opt-in limit)
    if (channels % 8 == 0 && per_ch * 8 <= 100 * 1024) {
        const size_t sm = per_ch * 8;
        if (sm > 48 * 1024) HVB_CUDA(cudaFuncSetAttribute(sppf_pool_concat_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        // float8 is not a CUDA vector type: Vec<8> below
        sppf_pool_concat_kernel<8><<<dim3(channels / 8, n), 256, sm, ctx->stream>>>(y0_dev, h, w, channels, out_cat_dev);
    } else {
        const size_t sm = per_ch * 4;
        if (sm > 48 * 1024) HVB_CUDA(cudaFuncSetAttribute(sppf_pool_concat_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        sppf_pool_concat_kernel<4><<<dim3(channels / 4, n), 256, sm, ctx->stream>>>(y0_dev, h, w, channels, out_cat_dev);
    }
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_stem_conv(hvb_ctx* ctx, const float* in_nchw_dev, const float* weight_host, const float* bias_host, int n, int h,
                  int w, int c_out, float* out_nhwc_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(in_nchw_dev && weight_host && out_nhwc_dev && n >= 0 && h > 0 && w > 0, "bad arguments");
    if (n == 0) return HVB_OK;
    HVB_ARG(n <= 65535 && (h + 1) / 2 <= 65535 * 4, "grid extent");
    switch (c_out) {
        case 16: return launch_stem<16>(ctx, in_nchw_dev, weight_host, bias_host, n, h, w, out_nhwc_dev);
        case 32: return launch_stem<32>(ctx, in_nchw_dev, weight_host, bias_host, n, h, w, out_nhwc_dev);
        case 48: return launch_stem<48>(ctx, in_nchw_dev, weight_host, bias_host, n, h, w, out_nhwc_dev);
        case 64: return launch_stem<64>(ctx, in_nchw_dev, weight_host, bias_host, n, h, w, out_nhwc_dev);
        default: hvb_set_error("hvb_stem_conv: c_out must be 16/32/48/64 (YOLOv8 n/s/m/l), got %d", c_out); return HVB_ERR_UNSUPPORTED;
    }
}

}  // extern "C"

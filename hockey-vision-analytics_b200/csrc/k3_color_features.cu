// K3a — per-detection jersey ROI -> BGR->HSV / BGR->LAB (bit-exact with cv2.cvtColor) ->
// histograms + exact integer moments -> the 49-d colour feature of
// HybridTeamClassifier.extract_color_features (reference hockey/common/team_hybrid.py:89-142).
//
// Work decomposition: a persistent grid of CTAs; each CTA copies the four lookup tables (8.5 KB)
// into shared memory once and then walks crops blockIdx.x, +gridDim.x, ...  Inside a crop every
// thread keeps PRIVATE packed 8-bit histogram counters and 32-bit moment sums in registers for a
// batch of <= 255 pixels, then the warp aggregates them with REDUX (one instruction per
// quantity) and a single lane per quantity adds the warp total into the CTA's shared-memory
// accumulators — no per-pixel atomics, so flat-colour ROIs (single-bin contention) cost the same
// as uniform-random ones.
#include "hvb_common.cuh"
#include "hvb_roi.cuh"

namespace {

constexpr int kThreads = 256;         // 8 warps per crop.  128 / 64 threads per crop with 8 / 16 CTAs per SM (every crop resident at once) measured
                                      // slower: K3a 29.1 / 35.4 us against 27.7 us for 768 crops, K3c 18.8 / 26.6 against 17.6 us (run r02z)
constexpr int kUnroll = 4;            // pixels whose loads are in flight per thread before the first one is used
constexpr int kBatch = 255;          // pixels per thread between flushes (8-bit packed counters)
constexpr int kNumU32 = 34 + 3;      // hist + counts
constexpr int kNumU64 = 12;          // sums + sums of squares

struct ColorTables {
    int32_t sdiv[256];
    int32_t hdiv[256];
    uint16_t gtab[256];
    uint16_t ctab[3072];
};
static_assert(sizeof(ColorTables) == HVB_TAB_BYTES, "table layout");
static_assert(sizeof(hvb_jersey_raw) == 120 && sizeof(hvb_color_raw) == 272, "raw struct layout");

__device__ __forceinline__ void load_tables(ColorTables* dst, const uint8_t* src_dev) {
    const uint4* s = reinterpret_cast<const uint4*>(src_dev);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < HVB_TAB_BYTES / 16; i += blockDim.x) d[i] = __ldg(s + i);
}

// OpenCV RGB2HSV_b, hsv_shift = 12, H range 180.
__device__ __forceinline__ void bgr_to_hsv(const ColorTables& t, int b, int g, int r, int& h, int& s, int& v) {
    v = max(max(b, g), r);
    int vmin = min(min(b, g), r);
    int diff = v - vmin;
    s = (diff * t.sdiv[v] + (1 << 11)) >> 12;
    int hh = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff);
    hh = (hh * t.hdiv[diff] + (1 << 11)) >> 12;   // arithmetic shift on a signed value
    h = hh < 0 ? hh + 180 : hh;
}

// OpenCV RGB2Lab_b, lab_shift = 12, lab_shift2 = 15.
__device__ __forceinline__ void bgr_to_lab(const ColorTables& t, int b, int g, int r, int& L, int& A, int& B) {
    int R = t.gtab[r], G = t.gtab[g], Bl = t.gtab[b];
    int fX = t.ctab[(R * 1777 + G * 1541 + Bl * 778 + (1 << 11)) >> 12];
    int fY = t.ctab[(R * 871 + G * 2929 + Bl * 296 + (1 << 11)) >> 12];
    int fZ = t.ctab[(R * 73 + G * 448 + Bl * 3575 + (1 << 11)) >> 12];
    int l = (296 * fY - 1336934 + (1 << 14)) >> 15;
    int a = (500 * (fX - fY) + 128 * 32768 + (1 << 14)) >> 15;
    int bb = (200 * (fY - fZ) + 128 * 32768 + (1 << 14)) >> 15;
    L = min(max(l, 0), 255);
    A = min(max(a, 0), 255);
    B = min(max(bb, 0), 255);
}

// Histograms: one private 32-bit counter per (bin, thread) in shared memory — address (bin * 256 + tid): every lane its
// own bank whatever the bin, so a flat-colour ROI (all lanes in one bin) costs the same as a random one, and an increment
// is one address multiply-add + one shared-memory reduction.  (Round 1 kept packed 8-bit counters in registers: ~35
// instructions per pixel for the three increments; measured 30.1 us against 28.8 us for 768 crops, run r02r.)
// The same two conversions with the tables addressed through their 32-bit shared-window address (formed once per CTA):
// through a generic reference the compiler re-derived the window base — S2R SR_CgaCtaId + LEA, twice — for every pixel.
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_u16(uint32_t a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return (int)v; }

__device__ __forceinline__ void bgr_to_hsv(uint32_t tab_sa, int b, int g, int r, int& h, int& s, int& v) {
    v = max(max(b, g), r);
    const int vmin = min(min(b, g), r);
    const int diff = v - vmin;
    s = (diff * lds_s32(tab_sa + HVB_TAB_SDIV + 4 * v) + (1 << 11)) >> 12;
    int hh = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff);
    hh = (hh * lds_s32(tab_sa + HVB_TAB_HDIV + 4 * diff) + (1 << 11)) >> 12;
    h = hh < 0 ? hh + 180 : hh;
}

__device__ __forceinline__ void bgr_to_lab(uint32_t tab_sa, int b, int g, int r, int& L, int& A, int& B) {
    const int R = lds_u16(tab_sa + HVB_TAB_GTAB + 2 * r), G = lds_u16(tab_sa + HVB_TAB_GTAB + 2 * g), Bl = lds_u16(tab_sa + HVB_TAB_GTAB + 2 * b);
    const int fX = lds_u16(tab_sa + HVB_TAB_CTAB + 2 * ((R * 1777 + G * 1541 + Bl * 778 + (1 << 11)) >> 12));
    const int fY = lds_u16(tab_sa + HVB_TAB_CTAB + 2 * ((R * 871 + G * 2929 + Bl * 296 + (1 << 11)) >> 12));
    const int fZ = lds_u16(tab_sa + HVB_TAB_CTAB + 2 * ((R * 73 + G * 448 + Bl * 3575 + (1 << 11)) >> 12));
    const int l = (296 * fY - 1336934 + (1 << 14)) >> 15;
    const int a = (500 * (fX - fY) + 128 * 32768 + (1 << 14)) >> 15;
    const int bb = (200 * (fY - fZ) + 128 * 32768 + (1 << 14)) >> 15;
    L = min(max(l, 0), 255);
    A = min(max(a, 0), 255);
    B = min(max(bb, 0), 255);
}

struct ThreadAcc {
    uint32_t cnt;                               // packed 8-bit: S<30, S>100, white
    uint32_t sum[6];
    uint32_t sq[6];
    __device__ __forceinline__ void clear() {
        cnt = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) { sum[i] = 0; sq[i] = 0; }
    }
};

__device__ __forceinline__ void hist_inc(uint32_t shared_addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(shared_addr) : "memory");
}

// priv_sa: shared-window address of this thread's counter of bin 0 (formed once per CTA; through a generic pointer the
// compiler re-derived the window base — two S2R and two LEA — for every pixel)
__device__ __forceinline__ void accumulate_pixel(uint32_t tab_sa, ThreadAcc& a, uint32_t priv_sa, int b, int g, int r) {
    int h, s, v, L, A, B;
    bgr_to_hsv(tab_sa, b, g, r, h, s, v);
    bgr_to_lab(tab_sa, b, g, r, L, A, B);
    const int hb = (h * 205) >> 11;                // h / 10 for 0 <= h < 180
    hist_inc(priv_sa + (uint32_t)hb * (kThreads * 4));
    hist_inc(priv_sa + (uint32_t)(18 + (s >> 5)) * (kThreads * 4));
    hist_inc(priv_sa + (uint32_t)(26 + (v >> 5)) * (kThreads * 4));
    a.cnt += (s < 30 ? 1u : 0u) | (s > 100 ? (1u << 8) : 0u) | ((v > 200 && s < 30) ? (1u << 16) : 0u);
    a.sum[0] += h; a.sum[1] += s; a.sum[2] += v; a.sum[3] += L; a.sum[4] += A; a.sum[5] += B;
    a.sq[0] += h * h; a.sq[1] += s * s; a.sq[2] += v * v; a.sq[3] += L * L; a.sq[4] += A * A; a.sq[5] += B * B;
}

// Warp-aggregate the per-thread accumulators and add them into the CTA accumulators.
__device__ __forceinline__ void flush(ThreadAcc& a, uint32_t* __restrict__ priv, uint32_t* s_u32, unsigned long long* s_u64) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    uint32_t mine0 = 0, mine1 = 0;      // lane k keeps quantity k (first 32) and quantity 32+k
#pragma unroll
    for (int k = 0; k < 34; k++) {
        const uint32_t part = priv[k * kThreads];
        priv[k * kThreads] = 0;
        uint32_t tot = __reduce_add_sync(full, part);
        if (k < 32) { if (lane == k) mine0 = tot; } else { if (lane == k - 32) mine1 = tot; }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint32_t tot = __reduce_add_sync(full, (a.cnt >> (8 * k)) & 0xffu);
        if (lane == 2 + k) mine1 = tot;            // quantities 34..36
    }
    if (mine0) atomicAdd(&s_u32[lane], mine0);
    if (lane < 5 && mine1) atomicAdd(&s_u32[32 + lane], mine1);
    unsigned long long m64 = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        uint32_t ts = __reduce_add_sync(full, a.sum[k]);
        uint32_t tq = __reduce_add_sync(full, a.sq[k]);
        if (lane == k) m64 = ts;
        if (lane == 6 + k) m64 = tq;
    }
    if (lane < 12 && m64) atomicAdd(&s_u64[lane], m64);
    a.clear();
}

__device__ __forceinline__ double u128_to_double(unsigned __int128 x) {
    return (double)(unsigned long long)(x >> 64) * 18446744073709551616.0 + (double)(unsigned long long)x;
}

#ifndef HVB_K3A_MINBLOCKS
#define HVB_K3A_MINBLOCKS 4            /* CTAs per SM the register allocation allows: 4 = up to 64 registers (nothing re-materialised in the pixel loop); 5 = 48, 6 = 40 measured no faster (run r02j) */
#endif
__global__ void __launch_bounds__(kThreads, HVB_K3A_MINBLOCKS)
color_features_kernel(const uint8_t* __restrict__ pixels, const hvb_crop_desc* __restrict__ crops, int n,
                      int roi_mode, const uint8_t* __restrict__ tables_dev, double* __restrict__ out_feat,
                      int64_t feat_stride, hvb_color_raw* __restrict__ out_raw) {
    __shared__ __align__(16) ColorTables tab;
    __shared__ uint32_t s_u32[40];
    __shared__ unsigned long long s_u64[kNumU64];
    __shared__ uint32_t s_priv[34 * kThreads];                  // [bin][thread] private histogram counters (34 KB)
    uint32_t* priv = s_priv + threadIdx.x;
    const uint32_t priv_sa = (uint32_t)__cvta_generic_to_shared(priv);
    const uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(&tab);
#pragma unroll
    for (int k = 0; k < 34; k++) priv[k * kThreads] = 0;        // own column only: no barrier needed
    // the crop descriptor is requested BEFORE the tables: its latency hides behind the 8.5 KB table copy instead of
    // following it (ncu r02d: a quarter of the stall samples sat on the table copy / descriptor / first pixel loads, three
    // dependent global round trips at the head of a CTA that lives ~15 us)
    hvb_crop_desc cd_next = crops[min((int)blockIdx.x, n - 1)];
    load_tables(&tab, tables_dev);

    for (int ci = blockIdx.x; ci < n; ci += gridDim.x) {
        if (threadIdx.x < 40) s_u32[threadIdx.x] = 0;
        if (threadIdx.x < kNumU64) s_u64[threadIdx.x] = 0ull;
        __syncthreads();

        const hvb_crop_desc cd = cd_next;
        if (ci + (int)gridDim.x < n) cd_next = crops[ci + gridDim.x];
        const hvb_rect rc = hvb_roi_rect(cd.h, cd.w, roi_mode);
        const int rw = max(rc.right - rc.left, 0), rh = max(rc.bottom - rc.top, 0);
        const int npx = rw * rh;
        const uint8_t* base = pixels + cd.offset + (int64_t)rc.top * cd.pitch + (int64_t)rc.left * 3;

        ThreadAcc acc;
        acc.clear();
        // Uniform trip counts for the whole CTA: every thread runs n_iter iterations (tail lanes
        // skip the pixel), so the warp-collective flush is always reached converged.
        const int n_iter = (npx + kThreads - 1) / kThreads;
        // This thread's pixels are p = tid, tid + 256, ...: (row, column) walk incrementally — 256 = drow * rw + dcol — so
        // the per-pixel address costs a few adds instead of a float reciprocal multiply, two fix-ups and a 64-bit
        // multiply-add (SASS of the previous loop: 41 of 168 instructions per pixel were addressing and predication).
        const int rws = max(rw, 1);
        const int drow = kThreads / rws, dcol = kThreads - drow * rws;
        const int64_t step_bytes = (int64_t)drow * cd.pitch + dcol * 3;
        const int64_t wrap_bytes = (int64_t)cd.pitch - (int64_t)rws * 3;
        const int row0 = (int)threadIdx.x / rws;
        int col = (int)threadIdx.x - row0 * rws;
        const uint8_t* px = base + (int64_t)row0 * cd.pitch + col * 3;
        const int n_mine = (int)threadIdx.x < npx ? (npx - 1 - (int)threadIdx.x) / kThreads + 1 : 0;
        for (int it0 = 0; it0 < n_iter; it0 += kBatch) {
            const int it1 = min(it0 + kBatch, n_iter);
            // ncu (profiles/r02a_ncu_k3.md): the one-pixel-at-a-time loop spent 4.6 of every 5.9 stall cycles per issue on
            // the pixel loads (long scoreboard) — kUnroll pixels are loaded before the first is consumed
            for (int it = it0; it < it1; it += kUnroll) {
                int pb[kUnroll], pg[kUnroll], pr[kUnroll];
                bool ok[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; u++) {
                    const bool in_batch = (it + u) < it1;          // the last group of a 255-step batch is partial: the
                    ok[u] = in_batch && (it + u) < n_mine;         // next batch starts at it1, so the walk must stop there too
                    pb[u] = pg[u] = pr[u] = 0;
                    if (ok[u]) { pb[u] = __ldg(px); pg[u] = __ldg(px + 1); pr[u] = __ldg(px + 2); }
                    if (in_batch) {
                        px += step_bytes;
                        col += dcol;
                        if (col >= rws) { col -= rws; px += wrap_bytes; }
                    }
                }
#pragma unroll
                for (int u = 0; u < kUnroll; u++)
                    if (ok[u]) accumulate_pixel(tab_sa, acc, priv_sa, pb[u], pg[u], pr[u]);
            }
            __syncwarp();
            flush(acc, priv, s_u32, s_u64);
        }
        __syncthreads();

        // ---- epilogue: features in float64 exactly as numpy evaluates them
        const int f = threadIdx.x;
        if (out_feat && f < 49) {
            double val;
            const double dn = (double)npx;
            if (npx == 0) {
                val = __longlong_as_double(0x7ff8000000000000ll);
            } else if (f < 34) {
                int lo = f < 18 ? 0 : f < 26 ? 18 : 26, hi = f < 18 ? 18 : f < 26 ? 26 : 34;
                uint32_t tot = 0;
                for (int k = lo; k < hi; k++) tot += s_u32[k];
                float denom = __fadd_rn((float)tot, 1e-7f);                  // float32: hist.sum() + 1e-7
                val = (double)__fdiv_rn((float)s_u32[f], denom);
            } else if (f < 46) {
                int ch = (f < 40) ? (f - 34) % 3 : 3 + (f - 40) % 3;
                bool is_std = (f >= 37 && f < 40) || (f >= 43);
                unsigned long long sx = s_u64[ch], sxx = s_u64[6 + ch];
                if (!is_std) {
                    val = __ddiv_rn(__ddiv_rn((double)sx, dn), 255.0);
                } else {
                    unsigned __int128 num = (unsigned __int128)sxx * (unsigned)npx - (unsigned __int128)sx * sx;
                    double var = __ddiv_rn(u128_to_double(num), __dmul_rn(dn, dn));
                    val = __ddiv_rn(sqrt(var), 255.0);
                }
            } else {
                val = __ddiv_rn((double)s_u32[34 + (f - 46)], dn);
            }
            out_feat[(int64_t)ci * feat_stride + f] = val;
        }
        if (out_raw) {
            hvb_color_raw* o = out_raw + ci;
            if (f < 34) o->hist[f] = s_u32[f];
            if (f < 3) o->counts[f] = s_u32[34 + f];
            if (f < 6) { o->sums[f] = s_u64[f]; o->sumsq[f] = s_u64[6 + f]; }
            if (f == 0) {
                o->n = (uint32_t)npx;
                o->roi[0] = rc.top; o->roi[1] = rc.bottom; o->roi[2] = rc.left; o->roi[3] = rc.right;
                o->pad_[0] = o->pad_[1] = 0;
            }
        }
        __syncthreads();
    }
}

// ---- SegmentationTeamClassifier.extract_jersey_colors (team_segmentation.py:98-148) over a rectangular mask.
// Same structure as color_features_kernel: private packed counters for <= 255 pixels per thread, one REDUX per
// quantity per warp, then shared-memory totals.
struct JerseyAcc {
    unsigned long long h0, h1, h2;      // packed 8-bit hue bins of the non-white pixels
    uint32_t white, sat_col, sat_all, val_all;
    __device__ __forceinline__ void clear() { h0 = h1 = h2 = 0ull; white = sat_col = sat_all = val_all = 0; }
};

__device__ __forceinline__ void jersey_flush(JerseyAcc& a, uint32_t* s_u32, unsigned long long* s_u64) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 18; k++) {
        unsigned long long w = (k < 8) ? a.h0 : (k < 16) ? a.h1 : a.h2;
        uint32_t tot = __reduce_add_sync(full, (uint32_t)(w >> (8 * (k & 7))) & 0xffu);
        if (lane == k) mine = tot;
    }
    {
        uint32_t tot = __reduce_add_sync(full, a.white);
        if (lane == 18) mine = tot;
    }
    if (lane < 19 && mine) atomicAdd(&s_u32[lane], mine);
    uint32_t t0 = __reduce_add_sync(full, a.sat_col);
    uint32_t t1 = __reduce_add_sync(full, a.sat_all);
    uint32_t t2 = __reduce_add_sync(full, a.val_all);
    unsigned long long m64 = lane == 0 ? t0 : lane == 1 ? t1 : t2;
    if (lane < 3 && m64) atomicAdd(&s_u64[lane], m64);
    a.clear();
}

__global__ void __launch_bounds__(kThreads)
jersey_color_stats_kernel(const uint8_t* __restrict__ pixels, const hvb_crop_desc* __restrict__ crops, int n,
                          int roi_mode, const uint8_t* __restrict__ tables_dev, hvb_jersey_raw* __restrict__ out_raw) {
    __shared__ __align__(16) ColorTables tab;
    __shared__ uint32_t s_u32[32];
    __shared__ unsigned long long s_u64[4];
    load_tables(&tab, tables_dev);
    const uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(&tab);

    for (int ci = blockIdx.x; ci < n; ci += gridDim.x) {
        if (threadIdx.x < 32) s_u32[threadIdx.x] = 0;
        if (threadIdx.x < 4) s_u64[threadIdx.x] = 0ull;
        __syncthreads();

        const hvb_crop_desc cd = crops[ci];
        const hvb_rect rc = hvb_roi_rect(cd.h, cd.w, roi_mode);
        const int rw = max(rc.right - rc.left, 0), rh = max(rc.bottom - rc.top, 0);
        const int npx = rw * rh;
        const uint8_t* base = pixels + cd.offset + (int64_t)rc.top * cd.pitch + (int64_t)rc.left * 3;

        JerseyAcc acc;
        acc.clear();
        const int n_iter = (npx + kThreads - 1) / kThreads;          // uniform trip count, see color_features_kernel
        // incremental (row, column) walk of this thread's pixels tid, tid + 256, ...: see color_features_kernel
        const int rws = max(rw, 1);
        const int drow = kThreads / rws, dcol = kThreads - drow * rws;
        const int64_t step_bytes = (int64_t)drow * cd.pitch + dcol * 3;
        const int64_t wrap_bytes = (int64_t)cd.pitch - (int64_t)rws * 3;
        const int row0 = (int)threadIdx.x / rws;
        int col = (int)threadIdx.x - row0 * rws;
        const uint8_t* px = base + (int64_t)row0 * cd.pitch + col * 3;
        const int n_mine = (int)threadIdx.x < npx ? (npx - 1 - (int)threadIdx.x) / kThreads + 1 : 0;
        for (int it0 = 0; it0 < n_iter; it0 += kBatch) {
            const int it1 = min(it0 + kBatch, n_iter);
            for (int it = it0; it < it1; it += kUnroll) {
                int pb[kUnroll], pg[kUnroll], pr[kUnroll];
                bool ok[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; u++) {
                    const bool in_batch = (it + u) < it1;          // the last group of a 255-step batch is partial: the
                    ok[u] = in_batch && (it + u) < n_mine;         // next batch starts at it1, so the walk must stop there too
                    pb[u] = pg[u] = pr[u] = 0;
                    if (ok[u]) { pb[u] = __ldg(px); pg[u] = __ldg(px + 1); pr[u] = __ldg(px + 2); }
                    if (in_batch) {
                        px += step_bytes;
                        col += dcol;
                        if (col >= rws) { col -= rws; px += wrap_bytes; }
                    }
                }
#pragma unroll
                for (int u = 0; u < kUnroll; u++) {
                    if (!ok[u]) continue;
                    const int b = pb[u], g = pg[u], r = pr[u];
                    int h, s, v, L, A, B;
                    bgr_to_hsv(tab_sa, b, g, r, h, s, v);
                    bgr_to_lab(tab_sa, b, g, r, L, A, B);
                    // np.abs(a - 128) < 10 on uint8 arrays: a < 128 wraps to >= 128, so only 128..137 pass
                    const bool white = (L > 200) && (A >= 128 && A < 138) && (B >= 128 && B < 138);
                    if (white) {
                        acc.white += 1;
                    } else {
                        int hb = (h * 205) >> 11;                    // h / 10 for 0 <= h < 180
                        unsigned long long one = 1ull << ((hb & 7) * 8);
                        acc.h0 += (hb < 8) ? one : 0ull;
                        acc.h1 += (hb >= 8 && hb < 16) ? one : 0ull;
                        acc.h2 += (hb >= 16) ? one : 0ull;
                        acc.sat_col += s;
                    }
                    acc.sat_all += s;
                    acc.val_all += v;
                }
            }
            __syncwarp();
            jersey_flush(acc, s_u32, s_u64);
        }
        __syncthreads();

        const int f = threadIdx.x;
        hvb_jersey_raw* o = out_raw + ci;
        if (f < 18) o->hue_hist[f] = s_u32[f];
        if (f == 18) o->white = s_u32[18];
        if (f == 19) o->n = (uint32_t)npx;
        if (f == 20) o->sat_colored = s_u64[0];
        if (f == 21) o->sat_all = s_u64[1];
        if (f == 22) o->val_all = s_u64[2];
        if (f == 23) { o->roi[0] = rc.top; o->roi[1] = rc.bottom; o->roi[2] = rc.left; o->roi[3] = rc.right; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
cvt_hsv_lab_kernel(const uint8_t* __restrict__ bgr, int64_t n_px, const uint8_t* __restrict__ tables_dev,
                   uint8_t* __restrict__ out_hsv, uint8_t* __restrict__ out_lab) {
    __shared__ __align__(16) ColorTables tab;
    load_tables(&tab, tables_dev);
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += (int64_t)gridDim.x * blockDim.x) {
        int b = bgr[3 * p], g = bgr[3 * p + 1], r = bgr[3 * p + 2];
        int h, s, v, L, A, B;
        bgr_to_hsv(tab, b, g, r, h, s, v);
        bgr_to_lab(tab, b, g, r, L, A, B);
        if (out_hsv) { out_hsv[3 * p] = (uint8_t)h; out_hsv[3 * p + 1] = (uint8_t)s; out_hsv[3 * p + 2] = (uint8_t)v; }
        if (out_lab) { out_lab[3 * p] = (uint8_t)L; out_lab[3 * p + 1] = (uint8_t)A; out_lab[3 * p + 2] = (uint8_t)B; }
    }
}

// sv.crop_image: np.round (half-to-even) -> int, then numpy basic slicing frame[y0:y1, x0:x1].
__device__ __forceinline__ void numpy_slice(int lo, int hi, int dim, int& start, int& len) {
    if (lo < 0) lo += dim;
    if (hi < 0) hi += dim;
    lo = min(max(lo, 0), dim);
    hi = min(max(hi, 0), dim);
    start = lo;
    len = max(hi - lo, 0);
}

__global__ void crops_from_boxes_kernel(const float* __restrict__ xyxy, const int32_t* __restrict__ frame_idx, int n,
                                        int H, int W, hvb_crop_desc* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x0 = (int)rintf(xyxy[4 * i + 0]), y0 = (int)rintf(xyxy[4 * i + 1]);
    int x1 = (int)rintf(xyxy[4 * i + 2]), y1 = (int)rintf(xyxy[4 * i + 3]);
    int xs, xl, ys, yl;
    numpy_slice(x0, x1, W, xs, xl);
    numpy_slice(y0, y1, H, ys, yl);
    int64_t fr = frame_idx ? frame_idx[i] : 0;
    hvb_crop_desc d;
    d.offset = (fr * H + ys) * (int64_t)W * 3 + (int64_t)xs * 3;
    d.pitch = W * 3;
    d.h = yl;
    d.w = xl;
    d.reserved = 0;
    out[i] = d;
}

}  // namespace

extern "C" {

int hvb_color_features(hvb_ctx* ctx, const uint8_t* pixels_dev, const hvb_crop_desc* crops_dev, int n, int roi_mode,
                       double* out_feat_dev, int64_t feat_row_stride, hvb_color_raw* out_raw_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0, "n < 0");
    HVB_ARG(roi_mode >= 0 && roi_mode <= 2, "bad roi_mode");
    HVB_ARG(!out_feat_dev || feat_row_stride >= 49, "feat_row_stride < 49");
    if (n == 0) return HVB_OK;
    HVB_ARG(pixels_dev && crops_dev, "null input");
    int grid = n < ctx->sm_count * 8 ? n : ctx->sm_count * 8;
    color_features_kernel<<<grid, kThreads, 0, ctx->stream>>>(pixels_dev, crops_dev, n, roi_mode,
                                                              (const uint8_t*)ctx->tables_dev, out_feat_dev,
                                                              feat_row_stride, out_raw_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_cvt_hsv_lab(hvb_ctx* ctx, const uint8_t* bgr_dev, int64_t n_px, uint8_t* out_hsv_dev, uint8_t* out_lab_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n_px >= 0, "n_px < 0");
    if (n_px == 0) return HVB_OK;
    HVB_ARG(bgr_dev != nullptr, "null input");
    int64_t blocks = (n_px + 255) / 256;
    int grid = (int)(blocks < (int64_t)ctx->sm_count * 8 ? blocks : (int64_t)ctx->sm_count * 8);
    cvt_hsv_lab_kernel<<<grid, 256, 0, ctx->stream>>>(bgr_dev, n_px, (const uint8_t*)ctx->tables_dev, out_hsv_dev,
                                                      out_lab_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_crops_from_boxes(hvb_ctx* ctx, const float* xyxy_dev, const int32_t* frame_idx_dev, int n, int frame_h,
                         int frame_w, hvb_crop_desc* out_crops_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0 && frame_h > 0 && frame_w > 0, "bad sizes");
    if (n == 0) return HVB_OK;
    HVB_ARG(xyxy_dev && out_crops_dev, "null pointer");
    crops_from_boxes_kernel<<<hvb_div_up(n, 128), 128, 0, ctx->stream>>>(xyxy_dev, frame_idx_dev, n, frame_h, frame_w,
                                                                          out_crops_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_color_features_host(hvb_ctx* ctx, const uint8_t* pixels_host, size_t pixel_bytes,
                            const hvb_crop_desc* crops_host, int n, int roi_mode, double* out_feat_host,
                            hvb_color_raw* out_raw_host) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0, "n < 0");
    if (n == 0) return HVB_OK;
    HVB_ARG(pixels_host && crops_host && out_feat_host, "null pointer");
    size_t off_crops = (pixel_bytes + 255) & ~(size_t)255;
    size_t off_feat = off_crops + (((size_t)n * sizeof(hvb_crop_desc) + 255) & ~(size_t)255);
    size_t off_raw = off_feat + (((size_t)n * 49 * sizeof(double) + 255) & ~(size_t)255);
    size_t total = off_raw + (size_t)n * sizeof(hvb_color_raw);
    uint8_t* d = nullptr;
    HVB_TRY(hvb_scratch(ctx, total, (void**)&d));
    HVB_CUDA(cudaMemcpyAsync(d, pixels_host, pixel_bytes, cudaMemcpyHostToDevice, ctx->stream));
    HVB_CUDA(cudaMemcpyAsync(d + off_crops, crops_host, (size_t)n * sizeof(hvb_crop_desc), cudaMemcpyHostToDevice,
                             ctx->stream));
    HVB_TRY(hvb_color_features(ctx, d, (const hvb_crop_desc*)(d + off_crops), n, roi_mode, (double*)(d + off_feat), 49,
                               out_raw_host ? (hvb_color_raw*)(d + off_raw) : nullptr));
    HVB_CUDA(cudaMemcpyAsync(out_feat_host, d + off_feat, (size_t)n * 49 * sizeof(double), cudaMemcpyDeviceToHost,
                             ctx->stream));
    if (out_raw_host)
        HVB_CUDA(cudaMemcpyAsync(out_raw_host, d + off_raw, (size_t)n * sizeof(hvb_color_raw), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    HVB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HVB_OK;
}

int hvb_jersey_color_stats(hvb_ctx* ctx, const uint8_t* pixels_dev, const hvb_crop_desc* crops_dev, int n, int roi_mode,
                           hvb_jersey_raw* out_raw_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0, "n < 0");
    HVB_ARG(roi_mode >= 0 && roi_mode <= 3, "bad roi_mode");
    if (n == 0) return HVB_OK;
    HVB_ARG(pixels_dev && crops_dev && out_raw_dev, "null pointer");
    int grid = n < ctx->sm_count * 8 ? n : ctx->sm_count * 8;
    jersey_color_stats_kernel<<<grid, kThreads, 0, ctx->stream>>>(pixels_dev, crops_dev, n, roi_mode,
                                                                  (const uint8_t*)ctx->tables_dev, out_raw_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_jersey_color_stats_host(hvb_ctx* ctx, const uint8_t* pixels_host, size_t pixel_bytes,
                                const hvb_crop_desc* crops_host, int n, int roi_mode, hvb_jersey_raw* out_raw_host) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0, "n < 0");
    if (n == 0) return HVB_OK;
    HVB_ARG(pixels_host && crops_host && out_raw_host, "null pointer");
    size_t off_crops = (pixel_bytes + 255) & ~(size_t)255;
    size_t off_raw = off_crops + (((size_t)n * sizeof(hvb_crop_desc) + 255) & ~(size_t)255);
    size_t total = off_raw + (size_t)n * sizeof(hvb_jersey_raw);
    uint8_t* d = nullptr;
    HVB_TRY(hvb_scratch(ctx, total, (void**)&d));
    HVB_CUDA(cudaMemcpyAsync(d, pixels_host, pixel_bytes, cudaMemcpyHostToDevice, ctx->stream));
    HVB_CUDA(cudaMemcpyAsync(d + off_crops, crops_host, (size_t)n * sizeof(hvb_crop_desc), cudaMemcpyHostToDevice,
                             ctx->stream));
    HVB_TRY(hvb_jersey_color_stats(ctx, d, (const hvb_crop_desc*)(d + off_crops), n, roi_mode,
                                   (hvb_jersey_raw*)(d + off_raw)));
    HVB_CUDA(cudaMemcpyAsync(out_raw_host, d + off_raw, (size_t)n * sizeof(hvb_jersey_raw), cudaMemcpyDeviceToHost,
                             ctx->stream));
    HVB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HVB_OK;
}

}  // extern "C"

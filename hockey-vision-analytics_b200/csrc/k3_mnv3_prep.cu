// K3b — MobileNetV3 crop preprocessing: self.preprocess of hockey/common/team_hybrid.py:31-36
// (ToPILImage -> Resize((128,64)) -> ToTensor -> Normalize) applied to the jersey ROI of every
// crop (team_hybrid.py:49-64, 73-77), batched into one float32[n,3,128,64] tensor.
//
// The resize is Pillow's ImagingResample with the bilinear (triangle) filter, reproduced bit-exactly
// (SURVEY.md App. A4): per-axis coefficients computed in float64 exactly like precompute_coeffs,
// normalised, converted to 22-bit fixed point, horizontal pass to a uint8 intermediate, then the
// vertical pass (order swapped when in_h > 100*in_w, as the installed Pillow does).  ToTensor is a
// true float32 /255 and Normalize a true (x-mean)/std (IEEE division); the BGR crop goes through
// the RGB statistics unswapped, like the reference.
//
// One CTA per crop.  Coefficients and the uint8 intermediate live in shared memory; tall ROIs are
// processed in output-row blocks so that the intermediate window always fits.  Two kernels share the
// work: the FAST one (<= 8 taps per sample, i.e. ROI up to 192 x 384 px — every player crop of a
// 1080p/4K rink frame) first stages the block's source rows into shared memory with aligned 32-bit
// loads, all of them in flight at once (the ROI rows are unaligned 3-byte pixels; the first/last
// partial word of a row is assembled from byte loads so nothing outside the ROI is touched), then runs
// both filter passes out of shared memory; ~50 KB per CTA, four CTAs resident per SM.  The GENERAL one
// (<= 48 taps, 87 KB, filters straight from global memory) only touches the crops the fast one skipped.
// ToTensor + Normalize are a 3 x 256-entry table built once per CTA with the exact float32 divisions, so
// the store loop does one shared-memory lookup per value.
#include "hvb_common.cuh"
#include "hvb_roi.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kOutW = 64, kOutH = 128;
constexpr int kPrecision = 22;

// KMAX: max taps per output sample; ROWS: rows of the uint8 intermediate held in shared memory
template <int KMAX, int ROWS>
struct SmemT {
    int bh[kOutW][2];                   // horizontal bounds: xmin, n
    int bv[kOutH][2];
    int kh[kOutW * KMAX];
    int kv[kOutH * KMAX];
    float lut[3][256];                  // ((v / 255) - mean[c]) / std[c], exact float32 divisions
    uint8_t inter[ROWS * kOutW * 3];
    int blk[4];
};
constexpr int kFastK = 8, kFastRows = 96;         // ROI up to 192 x 384 px (scale <= 3 per axis)
constexpr int kStageBytes = 20480;                // staged source rows of one row block (fast kernel)
constexpr int kGenK = 48, kGenRows = 256;         // ROI up to ~1400 x 2900 px

// Pillow precompute_coeffs + normalize_coeffs_8bpc for output sample `o` (bilinear, support 1); kk = the sample's own row
// of `ksize` taps.
__device__ void pil_coeffs(int in_size, int out_size, int o, int ksize, int* bounds, int* kk) {
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;         // 1.0 * filterscale
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn(__dadd_rn((double)o, 0.5), scale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < n; x++) {
        double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (t < 0.0) t = -t;
        double w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
        ww = __dadd_rn(ww, w);
    }
    for (int x = 0; x < n; x++) {
        double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (t < 0.0) t = -t;
        double w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
        if (ww != 0.0) w = __ddiv_rn(w, ww);
        kk[x] = (int)__dadd_rn(0.5, __dmul_rn(w, (double)(1 << kPrecision)));
    }
    for (int x = n; x < ksize; x++) kk[x] = 0;
    bounds[0] = xmin;
    bounds[1] = n;
}

__device__ __forceinline__ int pil_ksize(int in_size, int out_size) {
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double support = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(support) * 2 + 1;
}

__device__ __forceinline__ int clip8(int v) {
    v >>= kPrecision;
    return min(max(v, 0), 255);
}

__device__ __forceinline__ void write_out(const float (*lut)[256], float* __restrict__ out, uint8_t* __restrict__ out_u8, int y, int x, int v0, int v1, int v2) {
    const int v[3] = {v0, v1, v2};
#pragma unroll
    for (int c = 0; c < 3; c++) out[(c * kOutH + y) * kOutW + x] = lut[c][v[c]];
    if (out_u8) {
        uint8_t* o = out_u8 + (y * kOutW + x) * 3;
        o[0] = (uint8_t)v0; o[1] = (uint8_t)v1; o[2] = (uint8_t)v2;
    }
}

// ---- fast path, second version (round 2).  What ncu said about the first (profiles/r01c_ncu_k3_k5.md, r02a_ncu_k3.md):
// 5.6 k instructions per thread per crop — 1.5 k of them the float64 Pillow coefficients recomputed for every crop,
// the rest inflated by per-tap coefficient / bound loads from shared memory, per-item index divisions, byte-wide
// intermediate traffic and one 32-bit store per output value; issue slots 54 % busy with barrier and fixed-latency
// stalls on top (a serial row-block planner ran on thread 0 while 255 threads waited).  This version
//   * reads the coefficients from a table built ONCE per context for every source size the fast path accepts
//     (hvb_k3b_build_tables: 384 sizes x (64 + 128) samples x 8 taps, 2.9 MB, L2-resident),
//   * gives every thread a fixed output column in the horizontal pass, so its bounds and taps live in registers,
//   * keeps the uint8 intermediate as one packed 32-bit word per pixel (one shared-memory load per tap and pixel in the
//     vertical pass instead of three, one store per pixel in the horizontal pass instead of three),
//   * lets a thread produce four adjacent output pixels in the vertical pass: one 128-bit intermediate load per tap
//     and three 128-bit global stores per item instead of twelve 32-bit ones,
//   * plans the row blocks with one __syncthreads_count.
// The integer arithmetic (22-bit coefficients, rounding, uint8 round trip between the passes) is unchanged, so the
// output stays bit-identical to Pillow / torchvision (tests/test_gpu_mnv3.py).
constexpr int kTabStride = (kOutW + kOutH) * (2 + kFastK);      // ints per source size: H part (640) then V part (1280)
constexpr int kTabHInts = kOutW * (2 + kFastK);
constexpr int kTabSizes = 384;                                  // source sizes 1 .. 384

struct SmemFast {
    int vtab[kOutH * (2 + kFastK)];     // ymin[128], n[128], k[128][8] of this crop's source height
    float lut[3][256];
    uint32_t inter[kFastRows * kOutW];  // horizontal-pass result, b | g << 8 | r << 16 per pixel
    uint32_t roi[kStageBytes / 4];      // staged source rows: row r at byte r * pitch_s, ROI byte 0 at + shift_r
};

__device__ __forceinline__ bool mnv3_fast_case(int rw, int rh, int ksh, int ksv) {
    // empty ROIs, and ROIs within the tap budget that resize horizontally first (rh <= 100 * rw)
    return rw <= 0 || rh <= 0 || (ksh <= kFastK && ksv <= kFastK && !(rh > 100 * rw));
}

// One thread per (source size, output sample): the table rows the fast kernel loads instead of calling pil_coeffs.
__global__ void __launch_bounds__(kOutW + kOutH)
k3b_table_kernel(int* __restrict__ tab) {
    const int s = blockIdx.x + 1;
    const bool horiz = threadIdx.x < kOutW;
    const int o = horiz ? threadIdx.x : threadIdx.x - kOutW;
    const int out_size = horiz ? kOutW : kOutH;
    int* part = tab + (int64_t)blockIdx.x * kTabStride + (horiz ? 0 : kTabHInts);
    int bounds[2] = {0, 0};
    int kk[kFastK];
#pragma unroll
    for (int x = 0; x < kFastK; x++) kk[x] = 0;
    if (pil_ksize(s, out_size) <= kFastK) pil_coeffs(s, out_size, o, kFastK, bounds, kk);     // larger sizes never reach the fast kernel
    part[o] = bounds[0];
    part[out_size + o] = bounds[1];
#pragma unroll
    for (int x = 0; x < kFastK; x++) part[2 * out_size + o * kFastK + x] = kk[x];
}

// Bilinear taps are non-negative and sum to at most 2^22 + 4 (eight taps, each rounded up by at most 1/2): the rounded
// accumulator 2^21 + sum(v * k) lies in [0, 256 * 2^22), so Pillow's clip8 reduces to the shift — no min / max needed.
__device__ __forceinline__ uint32_t shr8(int v) { return (uint32_t)v >> kPrecision; }

template <int TAPS>
__device__ __forceinline__ void hpass_item(const uint8_t* __restrict__ src, int cnt, const int (&kk)[kFastK], uint32_t* __restrict__ dst) {
    int s0 = 1 << (kPrecision - 1), s1 = s0, s2 = s0;
#pragma unroll
    for (int x = 0; x < TAPS; x++)
        if (x < cnt) { s0 += src[3 * x] * kk[x]; s1 += src[3 * x + 1] * kk[x]; s2 += src[3 * x + 2] * kk[x]; }
    *dst = shr8(s0) | (shr8(s1) << 8) | (shr8(s2) << 16);
}

__global__ void __launch_bounds__(kThreads, 4)
mnv3_prep_fast_kernel(const uint8_t* __restrict__ pixels, const hvb_crop_desc* __restrict__ crops, int n, int roi_mode,
                      const int* __restrict__ tab, float* __restrict__ out, uint8_t* __restrict__ out_u8,
                      uint8_t* __restrict__ out_valid) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SmemFast& S = *reinterpret_cast<SmemFast*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int* v_ymin = S.vtab;
    const int* v_n = S.vtab + kOutH;
    const int* v_k = S.vtab + 2 * kOutH;

    bool lut_ready = false;
    // Work item = (crop, band of 64 output rows): with one crop per item the persistent grid (4 CTAs x 148 SMs = 592) ran
    // 768 crops as "1 or 2 crops per CTA" — a makespan of two crops for 1.3 crops of average work (ncu r02d: the SMs were
    // active for 63 % of the kernel's duration); half-crop items make it 1.5.
    for (int item = blockIdx.x; item < 2 * n; item += gridDim.x) {
        const int ci = item >> 1, band = item & 1;
        const int y_begin = band * (kOutH / 2), y_end = y_begin + kOutH / 2;
        const hvb_crop_desc cd = crops[ci];
        const hvb_rect rc = hvb_roi_rect(cd.h, cd.w, roi_mode);
        const int rw = max(rc.right - rc.left, 0), rh = max(rc.bottom - rc.top, 0);
        const int ksh = (rw > 0) ? pil_ksize(rw, kOutW) : 1, ksv = (rh > 0) ? pil_ksize(rh, kOutH) : 1;
        if (!mnv3_fast_case(rw, rh, ksh, ksv)) continue;            // the general kernel owns this crop
        float* o = out + (int64_t)ci * 3 * kOutH * kOutW;
        uint8_t* o8 = out_u8 ? out_u8 + (int64_t)ci * kOutH * kOutW * 3 : nullptr;
        if (rw <= 0 || rh <= 0) {
            // empty ROI: the reference's `except:` path (zero feature row)
            for (int c = 0; c < 3; c++)
                for (int i = threadIdx.x; i < (kOutH / 2) * kOutW; i += kThreads) o[(c * kOutH + y_begin) * kOutW + i] = 0.f;
            if (o8) for (int i = threadIdx.x; i < 3 * (kOutH / 2) * kOutW; i += kThreads) o8[3 * y_begin * kOutW + i] = 0;
            if (out_valid && threadIdx.x == 0 && band == 0) out_valid[ci] = 0;
            continue;
        }
        if (!lut_ready) {
            const float mean[3] = {0.485f, 0.456f, 0.406f};
            const float stdv[3] = {0.229f, 0.224f, 0.225f};
            for (int i = threadIdx.x; i < 3 * 256; i += kThreads) {
                const int c = i >> 8, v = i & 255;
                S.lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean[c]), stdv[c]);   // ToTensor, Normalize
            }
            lut_ready = true;                       // visible after the __syncthreads below
        }
        if (out_valid && threadIdx.x == 0 && band == 0) out_valid[ci] = 1;
        const uint8_t* base = pixels + cd.offset + (int64_t)rc.top * cd.pitch + (int64_t)rc.left * 3;
        const int row_bytes = rw * 3;
        const int pitch_s = ((row_bytes + 3 + 3) >> 2) << 2;       // room for a shift of up to 3 bytes, word multiple
        const int rmax = min(kFastRows, kStageBytes / pitch_s);    // >= 8 = kFastK source rows always fit

        __syncthreads();   // previous crop's readers are done with the tables
        {   // vertical table of this source height -> shared memory; horizontal taps of this thread's column -> registers
            const int4* tv = reinterpret_cast<const int4*>(tab + (int64_t)(rh - 1) * kTabStride + kTabHInts);
            int4* dv = reinterpret_cast<int4*>(S.vtab);
            for (int i = threadIdx.x; i < kOutH * (2 + kFastK) / 4; i += kThreads) dv[i] = __ldg(tv + i);
        }
        const int xx = threadIdx.x & (kOutW - 1), hq = threadIdx.x >> 6;
        const int* th = tab + (int64_t)(rw - 1) * kTabStride;
        const int hxmin = __ldg(th + xx), hcnt = __ldg(th + kOutW + xx);
        int hk[kFastK];
        {
            const int4 k0 = __ldg(reinterpret_cast<const int4*>(th + 2 * kOutW + xx * kFastK));
            const int4 k1 = __ldg(reinterpret_cast<const int4*>(th + 2 * kOutW + xx * kFastK) + 1);
            hk[0] = k0.x; hk[1] = k0.y; hk[2] = k0.z; hk[3] = k0.w; hk[4] = k1.x; hk[5] = k1.y; hk[6] = k1.z; hk[7] = k1.w;
        }
        __syncthreads();

        int y0 = y_begin;
        while (y0 < y_end) {
            // largest block of output rows whose source-row window fits the staging + intermediate buffers: the window
            // end ymin + n is non-decreasing in y, so the rows that fit are a prefix and one block-wide count finds it
            const int r_lo = v_ymin[y0];
            const int y = threadIdx.x;
            const int fits = __syncthreads_count(y >= y0 && y < y_end && v_ymin[y] + v_n[y] - r_lo <= rmax);
            const int y1 = y0 + max(fits, 1);
            const int r_hi = v_ymin[y1 - 1] + v_n[y1 - 1];
            const int nrows = r_hi - r_lo;
            // ---- stage rows [r_lo, r_hi): a warp takes FOUR rows per pass (r, r + 8, r + 16, r + 24), a lane two aligned
            // 32-bit words of each — up to eight independent loads in flight per lane before the first shared store.  ncu
            // on the one-row-at-a-time loop (profiles/r02d_ncu_k3.md, tools/ncu_lines.py): 30 % of the kernel's stall samples
            // sat on this phase (a 45-pixel row is 34 words: two dependent global-latency round trips per row and warp).
            {
                const int wpr = pitch_s >> 2;
                const int slack_lo = min(rc.left * 3, 3), slack_hi = min((cd.w - rc.right) * 3, 3);   // crop-row bytes beside the ROI
                for (int rb = warp; rb < nrows; rb += 4 * (kThreads / 32)) {
                    uint32_t v[4][2];
                    bool live[4][2];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int r = rb + u * (kThreads / 32);
                        const uint8_t* rowp = base + (int64_t)(r_lo + min(r, nrows - 1)) * cd.pitch;
                        const int sh = (int)(reinterpret_cast<uintptr_t>(rowp) & 3);
                        const uint8_t* ap = rowp - sh;                  // aligned-down address of the row's first word
                        const int nwords = (sh + row_bytes + 3) >> 2;
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const int i = lane + 32 * j;
                            live[u][j] = r < nrows && i < nwords;
                            v[u][j] = 0;
                            if (live[u][j]) {
                                const int b0 = 4 * i - sh;              // ROI byte index of the word's first byte
                                // whole words: inside the ROI row, or (partial first / last word) still inside the CROP's own
                                // row, whose bytes belong to the caller's buffer — the extra bytes are never read back
                                if (b0 >= -slack_lo && b0 + 4 <= row_bytes + slack_hi) {
                                    v[u][j] = __ldg(reinterpret_cast<const uint32_t*>(ap) + i);
                                } else {                                // at the edge of the crop row: only bytes of the ROI
#pragma unroll
                                    for (int k = 0; k < 4; k++)
                                        if (b0 + k >= 0 && b0 + k < row_bytes) v[u][j] |= (uint32_t)__ldg(rowp + b0 + k) << (8 * k);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int j = 0; j < 2; j++)
                            if (live[u][j]) S.roi[(rb + u * (kThreads / 32)) * wpr + lane + 32 * j] = v[u][j];
                }
                if (wpr > 64) {                                         // rows wider than 64 words (ROI wider than 84 px): the rest
                    for (int r = warp; r < nrows; r += kThreads / 32) {
                        const uint8_t* rowp = base + (int64_t)(r_lo + r) * cd.pitch;
                        const int sh = (int)(reinterpret_cast<uintptr_t>(rowp) & 3);
                        const uint8_t* ap = rowp - sh;
                        const int nwords = (sh + row_bytes + 3) >> 2;
                        for (int i = 64 + lane; i < nwords; i += 32) {
                            uint32_t w = 0;
                            const int b0 = 4 * i - sh;
                            if (b0 + 4 <= row_bytes) {
                                w = __ldg(reinterpret_cast<const uint32_t*>(ap) + i);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; k++)
                                    if (b0 + k < row_bytes) w |= (uint32_t)__ldg(rowp + b0 + k) << (8 * k);
                            }
                            S.roi[r * wpr + i] = w;
                        }
                    }
                }
            }
            __syncthreads();
            // ---- horizontal pass out of shared memory: thread = (column xx, row slot hq); rows hq, hq + 4, ...
            {
                const uint8_t* sroi = reinterpret_cast<const uint8_t*>(S.roi);
                const int sh0 = (int)(reinterpret_cast<uintptr_t>(base + (int64_t)r_lo * cd.pitch) & 3), dsh = cd.pitch & 3;
                if (ksh <= 3) {                                 // ROI not wider than 64 px (no horizontal down-scaling): 3 taps
                    for (int row = hq; row < nrows; row += kThreads / kOutW)
                        hpass_item<3>(sroi + row * pitch_s + ((sh0 + row * dsh) & 3) + hxmin * 3, hcnt, hk, S.inter + row * kOutW + xx);
                } else if (ksh <= 5) {                          // up to 128 px wide: 5 taps
                    for (int row = hq; row < nrows; row += kThreads / kOutW)
                        hpass_item<5>(sroi + row * pitch_s + ((sh0 + row * dsh) & 3) + hxmin * 3, hcnt, hk, S.inter + row * kOutW + xx);
                } else {
                    for (int row = hq; row < nrows; row += kThreads / kOutW)
                        hpass_item<kFastK>(sroi + row * pitch_s + ((sh0 + row * dsh) & 3) + hxmin * 3, hcnt, hk, S.inter + row * kOutW + xx);
                }
            }
            __syncthreads();
            // ---- vertical pass: thread = (four adjacent output columns 4g .. 4g + 3, row slot yq); rows y0 + yq, + 16, ...
            {
                const int g = threadIdx.x & 15, yq = threadIdx.x >> 4;
                for (int yy = y0 + yq; yy < y1; yy += kThreads / 16) {
                    const int ymin = v_ymin[yy], cnt = v_n[yy];
                    const int* k = v_k + yy * kFastK;
                    const uint4* src = reinterpret_cast<const uint4*>(S.inter + (ymin - r_lo) * kOutW) + g;
                    int a[4][3];
#pragma unroll
                    for (int px = 0; px < 4; px++) { a[px][0] = a[px][1] = a[px][2] = 1 << (kPrecision - 1); }
                    for (int t = 0; t < cnt; t++) {
                        const uint4 w4 = src[t * (kOutW / 4)];
                        const int kk = k[t];
                        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int px = 0; px < 4; px++) {
                            a[px][0] += (int)(w[px] & 0xffu) * kk;
                            a[px][1] += (int)((w[px] >> 8) & 0xffu) * kk;
                            a[px][2] += (int)(w[px] >> 16) * kk;
                        }
                    }
                    int v[4][3];
#pragma unroll
                    for (int px = 0; px < 4; px++) { v[px][0] = (int)shr8(a[px][0]); v[px][1] = (int)shr8(a[px][1]); v[px][2] = (int)shr8(a[px][2]); }
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float4 f;
                        f.x = S.lut[c][v[0][c]]; f.y = S.lut[c][v[1][c]]; f.z = S.lut[c][v[2][c]]; f.w = S.lut[c][v[3][c]];
                        *reinterpret_cast<float4*>(o + (c * kOutH + yy) * kOutW + 4 * g) = f;
                    }
                    if (o8) {
                        uint8_t* d8 = o8 + (yy * kOutW + 4 * g) * 3;
#pragma unroll
                        for (int px = 0; px < 4; px++) { d8[3 * px] = (uint8_t)v[px][0]; d8[3 * px + 1] = (uint8_t)v[px][1]; d8[3 * px + 2] = (uint8_t)v[px][2]; }
                    }
                }
            }
            __syncthreads();
            y0 = y1;
        }
    }
}

// FAST = true: handles the crops that fit the small tables and skips the rest; FAST = false: the complement.
template <int KMAX, int ROWS, bool FAST>
__global__ void __launch_bounds__(kThreads)
mnv3_prep_kernel(const uint8_t* __restrict__ pixels, const hvb_crop_desc* __restrict__ crops, int n, int roi_mode,
                 float* __restrict__ out, uint8_t* __restrict__ out_u8, uint8_t* __restrict__ out_valid) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    typedef SmemT<KMAX, ROWS> Smem;
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    constexpr int kKMax = KMAX, kInterRows = ROWS;

    bool lut_ready = false;
    for (int ci = blockIdx.x; ci < n; ci += gridDim.x) {
        const hvb_crop_desc cd = crops[ci];
        const hvb_rect rc = hvb_roi_rect(cd.h, cd.w, roi_mode);
        const int rw = max(rc.right - rc.left, 0), rh = max(rc.bottom - rc.top, 0);
        float* o = out + (int64_t)ci * 3 * kOutH * kOutW;
        uint8_t* o8 = out_u8 ? out_u8 + (int64_t)ci * kOutH * kOutW * 3 : nullptr;
        const int ksh = (rw > 0) ? pil_ksize(rw, kOutW) : 1, ksv = (rh > 0) ? pil_ksize(rh, kOutH) : 1;
        const bool vfirst = rh > 100 * rw;
        // the fast instantiation owns: empty ROIs, and ROIs within its tap budget that resize horizontally first
        if (mnv3_fast_case(rw, rh, ksh, ksv) != FAST) continue;
        if (!lut_ready) {
            const float mean[3] = {0.485f, 0.456f, 0.406f};
            const float stdv[3] = {0.229f, 0.224f, 0.225f};
            for (int i = threadIdx.x; i < 3 * 256; i += kThreads) {
                const int c = i >> 8, v = i & 255;
                S.lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean[c]), stdv[c]);   // ToTensor, Normalize
            }
            lut_ready = true;                       // visible after the __syncthreads below
        }
        const bool ok = rw > 0 && rh > 0 && ksh <= kKMax && ksv <= kKMax && (!vfirst || rw * kOutH * 3 <= (int)sizeof(S.inter));
        if (!ok) {
            // empty ROI: the reference's `except:` path (zero feature row); too-large ROI: flagged 0xFF
            for (int i = threadIdx.x; i < 3 * kOutH * kOutW; i += kThreads) o[i] = 0.f;
            if (o8) for (int i = threadIdx.x; i < 3 * kOutH * kOutW; i += kThreads) o8[i] = 0;
            if (out_valid && threadIdx.x == 0) out_valid[ci] = (rw > 0 && rh > 0) ? 0xFF : 0;
            continue;
        }
        if (out_valid && threadIdx.x == 0) out_valid[ci] = 1;
        const uint8_t* base = pixels + cd.offset + (int64_t)rc.top * cd.pitch + (int64_t)rc.left * 3;

        __syncthreads();   // previous crop's readers are done with the tables
        if (threadIdx.x < kOutW) pil_coeffs(rw, kOutW, threadIdx.x, ksh, S.bh[threadIdx.x], S.kh + threadIdx.x * ksh);
        else if (threadIdx.x >= 128 && threadIdx.x < 128 + kOutH) pil_coeffs(rh, kOutH, threadIdx.x - 128, ksv, S.bv[threadIdx.x - 128], S.kv + (threadIdx.x - 128) * ksv);
        __syncthreads();

        if (!vfirst) {
            int y0 = 0;
            while (y0 < kOutH) {
                // largest block of output rows whose source-row window fits the intermediate buffer
                if (threadIdx.x == 0) {
                    const int r_lo = S.bv[y0][0];
                    int y1 = y0 + 1;
                    while (y1 < kOutH && S.bv[y1][0] + S.bv[y1][1] - r_lo <= kInterRows) y1++;
                    S.blk[0] = y1; S.blk[1] = r_lo; S.blk[2] = S.bv[y1 - 1][0] + S.bv[y1 - 1][1];
                }
                __syncthreads();
                const int y1 = S.blk[0], r_lo = S.blk[1], r_hi = S.blk[2];
                // horizontal pass: rows [r_lo, r_hi) -> inter[row - r_lo][xx][c]
                for (int it = threadIdx.x; it < (r_hi - r_lo) * kOutW; it += kThreads) {
                    const int row = it / kOutW, xx = it % kOutW;
                    const int xmin = S.bh[xx][0], cnt = S.bh[xx][1];
                    const uint8_t* src = base + (int64_t)(r_lo + row) * cd.pitch + xmin * 3;
                    const int* k = S.kh + xx * ksh;
                    int s0 = 1 << (kPrecision - 1), s1 = s0, s2 = s0;
                    for (int x = 0; x < cnt; x++) {
                        const int kk = k[x];
                        s0 += __ldg(src + 3 * x) * kk; s1 += __ldg(src + 3 * x + 1) * kk; s2 += __ldg(src + 3 * x + 2) * kk;
                    }
                    uint8_t* d = S.inter + (row * kOutW + xx) * 3;
                    d[0] = (uint8_t)clip8(s0); d[1] = (uint8_t)clip8(s1); d[2] = (uint8_t)clip8(s2);
                }
                __syncthreads();
                // vertical pass: output rows [y0, y1)
                for (int it = threadIdx.x; it < (y1 - y0) * kOutW; it += kThreads) {
                    const int y = y0 + it / kOutW, x = it % kOutW;
                    const int ymin = S.bv[y][0], cnt = S.bv[y][1];
                    const int* k = S.kv + y * ksv;
                    const uint8_t* src = S.inter + ((ymin - r_lo) * kOutW + x) * 3;
                    int s0 = 1 << (kPrecision - 1), s1 = s0, s2 = s0;
                    for (int t = 0; t < cnt; t++) {
                        const int kk = k[t];
                        s0 += src[t * kOutW * 3] * kk; s1 += src[t * kOutW * 3 + 1] * kk; s2 += src[t * kOutW * 3 + 2] * kk;
                    }
                    write_out(S.lut, o, o8, y, x, clip8(s0), clip8(s1), clip8(s2));
                }
                __syncthreads();
                y0 = y1;
            }
        } else {
            // degenerate tall-narrow ROI: vertical pass first -> inter[y][col][c] (128 x rw), then horizontal
            for (int it = threadIdx.x; it < kOutH * rw; it += kThreads) {
                const int y = it / rw, col = it % rw;
                const int ymin = S.bv[y][0], cnt = S.bv[y][1];
                const int* k = S.kv + y * ksv;
                const uint8_t* src = base + (int64_t)ymin * cd.pitch + col * 3;
                int s0 = 1 << (kPrecision - 1), s1 = s0, s2 = s0;
                for (int t = 0; t < cnt; t++) {
                    const int kk = k[t];
                    const uint8_t* p = src + (int64_t)t * cd.pitch;
                    s0 += __ldg(p) * kk; s1 += __ldg(p + 1) * kk; s2 += __ldg(p + 2) * kk;
                }
                uint8_t* d = S.inter + (y * rw + col) * 3;
                d[0] = (uint8_t)clip8(s0); d[1] = (uint8_t)clip8(s1); d[2] = (uint8_t)clip8(s2);
            }
            __syncthreads();
            for (int it = threadIdx.x; it < kOutH * kOutW; it += kThreads) {
                const int y = it / kOutW, xx = it % kOutW;
                const int xmin = S.bh[xx][0], cnt = S.bh[xx][1];
                const int* k = S.kh + xx * ksh;
                const uint8_t* src = S.inter + (y * rw + xmin) * 3;
                int s0 = 1 << (kPrecision - 1), s1 = s0, s2 = s0;
                for (int x = 0; x < cnt; x++) {
                    const int kk = k[x];
                    s0 += src[3 * x] * kk; s1 += src[3 * x + 1] * kk; s2 += src[3 * x + 2] * kk;
                }
                write_out(S.lut, o, o8, y, xx, clip8(s0), clip8(s1), clip8(s2));
            }
            __syncthreads();
        }
    }
}

}  // namespace

// Called once from hvb_ctx_create: Pillow's coefficients for every source size the fast kernel accepts (2.9 MB).
int hvb_k3b_build_tables(hvb_ctx* ctx) {
    HVB_CUDA(cudaMalloc(&ctx->k3b_tab_dev, (size_t)kTabSizes * kTabStride * sizeof(int)));
    k3b_table_kernel<<<kTabSizes, kOutW + kOutH, 0, ctx->stream>>>((int*)ctx->k3b_tab_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

extern "C" {

int hvb_mnv3_preprocess(hvb_ctx* ctx, const uint8_t* pixels_dev, const hvb_crop_desc* crops_dev, int n, int roi_mode,
                        float* out_dev, uint8_t* out_u8_dev, uint8_t* out_valid_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0, "n < 0");
    HVB_ARG(roi_mode >= 0 && roi_mode <= 2, "bad roi_mode");
    if (n == 0) return HVB_OK;
    HVB_ARG(pixels_dev && crops_dev && out_dev, "null pointer");
    typedef SmemT<kGenK, kGenRows> SmemGen;
    static_assert(sizeof(SmemFast) <= 54 * 1024 && sizeof(SmemGen) < 200 * 1024, "smem budget");
    static_assert(kFastK == 8 && kTabSizes >= 384, "table layout");
    HVB_ARG(ctx->k3b_tab_dev != nullptr, "context has no K3b coefficient table");
    auto fast = mnv3_prep_fast_kernel;
    auto gen = mnv3_prep_kernel<kGenK, kGenRows, false>;
    HVB_CUDA(cudaFuncSetAttribute(fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemFast)));
    HVB_CUDA(cudaFuncSetAttribute(gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemGen)));
    // fast pass: items = (crop, half of the output rows); at most the four CTAs (64 registers x 256 threads) resident per SM
    const int gfast = 2 * n < ctx->sm_count * 4 ? 2 * n : ctx->sm_count * 4;
    fast<<<gfast, kThreads, sizeof(SmemFast), ctx->stream>>>(pixels_dev, crops_dev, n, roi_mode, (const int*)ctx->k3b_tab_dev, out_dev,
                                                             out_u8_dev, out_valid_dev);
    HVB_LAUNCHED(ctx);
    // general pass: only the crops the fast pass skipped (usually none: its CTAs read the descriptors and exit)
    const int ggen = n < ctx->sm_count ? n : ctx->sm_count;
    gen<<<ggen, kThreads, sizeof(SmemGen), ctx->stream>>>(pixels_dev, crops_dev, n, roi_mode, out_dev, out_u8_dev, out_valid_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_mnv3_preprocess_host(hvb_ctx* ctx, const uint8_t* pixels_host, size_t pixel_bytes, const hvb_crop_desc* crops_host,
                             int n, int roi_mode, float* out_host, uint8_t* out_valid_host) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0, "n < 0");
    if (n == 0) return HVB_OK;
    HVB_ARG(pixels_host && crops_host && out_host, "null pointer");
    const size_t per = (size_t)3 * kOutH * kOutW * sizeof(float);
    size_t o_crops = (pixel_bytes + 255) & ~(size_t)255;
    size_t o_out = o_crops + (((size_t)n * sizeof(hvb_crop_desc) + 255) & ~(size_t)255);
    size_t o_valid = o_out + (size_t)n * per;
    uint8_t* d = nullptr;
    HVB_TRY(hvb_scratch(ctx, o_valid + n, (void**)&d));
    HVB_CUDA(cudaMemcpyAsync(d, pixels_host, pixel_bytes, cudaMemcpyHostToDevice, ctx->stream));
    HVB_CUDA(cudaMemcpyAsync(d + o_crops, crops_host, (size_t)n * sizeof(hvb_crop_desc), cudaMemcpyHostToDevice, ctx->stream));
    HVB_TRY(hvb_mnv3_preprocess(ctx, d, (const hvb_crop_desc*)(d + o_crops), n, roi_mode, (float*)(d + o_out), nullptr,
                                d + o_valid));
    HVB_CUDA(cudaMemcpyAsync(out_host, d + o_out, (size_t)n * per, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_valid_host) HVB_CUDA(cudaMemcpyAsync(out_valid_host, d + o_valid, n, cudaMemcpyDeviceToHost, ctx->stream));
    HVB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HVB_OK;
}

}  // extern "C"

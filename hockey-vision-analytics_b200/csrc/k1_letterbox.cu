// K1a / K1b — BGR uint8 frames -> letterboxed, RGB, /255, NCHW float32 batches, for whole frames
// (ultralytics LetterBox + preprocess behind hockey/main.py:179-184) and for InferenceSlicer
// tiles (supervision _generate_offset + crop_image, then LetterBox per tile; README.md:25).
//
// The uint8 stage is bit-exact with cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder(114)
// (SURVEY.md App. A5): 11-bit fixed-point coefficients built in float32 exactly like OpenCV
// (on the host, once per geometry), horizontal fraction clamped at the borders, vertical indices
// clamped with the fraction kept, the (>>4, >>16, +2 >>2) vertical rounding, and the 2x2
// area-average shortcut when both scale factors are exactly 2.  The float stage reproduces
// numpy/torch `x / 255` in float32 exactly with ONE multiply: float(v) * 0x1.010102p-8 rounded
// toward zero equals float(v) / 255 (round-to-nearest) for every v in [0, 255] (checked
// exhaustively on the host and again by tests/test_gpu_letterbox.py).
//
// One launch covers a chunk of frames: grid = blocks_per_frame x n_frames.  Each CTA produces an
// 16-row x 128-column block of one tile's output.  It first stages the source pixels the block
// needs into shared memory AS 32-BIT PIXEL WORDS (B | G<<8 | R<<16): 4-pixel groups are fetched
// with three aligned 32-bit loads and re-packed with byte permutes into one 128-bit shared store,
// so that afterwards one LDS.32 fetches a whole pixel.  The 2-tap horizontal filter of a channel is
// then ONE dp2a (two 16-bit coefficients x two 8-bit samples) after a byte permute that pairs the
// samples of the two taps; every thread owns ONE output column and walks down the 16 rows of the block,
// reusing the horizontal result of a source row shared by consecutive output rows; a warp writes 128 contiguous bytes per
// plane per row.  The vertical pass and the byte -> float conversion are arranged so that their work sits on the FMA pipe
// (IMAD, IDP.2A, FFMA) rather than on the ALU pipe the shifts of OpenCV's rounding rule would otherwise saturate.
#include "hvb_common.cuh"

#include <algorithm>
#include <type_traits>
#include <math.h>
#include <stdlib.h>
#include <map>


namespace {

#ifndef HVB_K1_TH
#define HVB_K1_TH 16
#endif
#ifndef HVB_K1_TW
#define HVB_K1_TW 128           /* run r02x: 128 x 16 blocks against 256 x 16 — K1b 479 vs 523 us per 16 4K frames (a 640-wide tile is
                                   five 128-column blocks but two and a HALF 256-column ones), K1a 186 vs 189 us; 8 / 32 rows measured slower */
#endif
constexpr int kRowUnroll = 4;   // rows of the column loop unrolled together (2 / 16 measured the same, 8 spills: run r02za)
constexpr int kTH = HVB_K1_TH;  // output rows per CTA
constexpr int kTW = HVB_K1_TW;  // output columns per CTA
constexpr int kThreads = kTW;   // one thread per output column of the block

enum { MODE_COPY = 0, MODE_LINEAR = 1, MODE_AREA2 = 2 };

struct LbJob {
    int64_t src_off;     // byte offset of the tile's first source pixel inside frame 0
    int64_t out_off;     // element offset of this tile's output block for frame 0
    int64_t out_frame_stride;  // elements between the same tile of consecutive frames
    int32_t src_w, src_h, new_w, new_h, top, left, out_h, out_w;
    int32_t mode;
    int32_t xtab, ytab;  // offsets (entries) of this job's coefficient tables
    int32_t row_px;      // pixels from the tile origin to the end of the frame row
    int32_t vec_ok;      // 4-pixel groups of this tile are 4-byte aligned in global memory
    int32_t pad_[3];
};

struct __align__(16) LbBlock {     // one per CTA of a frame: which tile/rows/cols, and its source window (host-computed)
    uint32_t packed;               // job << 24 | block_row << 12 | block_col
    int32_t flags;                 // bit 0: block has source pixels; bit 1: aligned 4-pixel-group loads allowed
    int32_t gx_lo, sy_lo;          // first staged pixel (multiple of 4) and row, tile-relative
    int32_t n_rows, n_groups;      // staged rows and 4-pixel groups per row
    int32_t row_px;                // pixels from gx_lo to the end of the frame row
    int32_t pad_;
    int64_t src_off;               // byte offset of pixel (gx_lo, sy_lo) inside frame 0
    int64_t pad2_;
};                                 // 48 bytes: everything the staging phase needs, one dependent load

struct XCoef { int32_t sx; uint32_t apk; };              // apk = a0 | a1 << 16
struct __align__(16) YCoef { int32_t r0, r1; int32_t b0, b1; };

__device__ __forceinline__ float u8_over_255(int v) {
    return __fmul_rz(__int2float_rn(v), 0x1.010102p-8f);   // == float(v) / 255.0f for 0 <= v <= 255
}

// float(byte SEL of w) / 255 without an int->float conversion or a shift: a byte permute drops the byte into the
// mantissa of 2^23 (0x4B0000vv == 8388608 + v), and ONE fused multiply-add rounded toward zero removes the offset:
// (2^23 + v) * k - 2^23 * k == v * k exactly inside the FMA (2^23 * k is a power-of-two multiple of k, hence
// representable), so the result is bit for bit u8_over_255(v).  One ALU-pipe + one FMA-pipe instruction per value
// instead of an extract, an I2FP and an FMUL (the A/B of the pieces is in profiles/r02w_k1_variants.md).
template <int SEL>
__device__ __forceinline__ float byte_over_255(uint32_t w) {
    const uint32_t bits = __byte_perm(w, 0x4B000000u, 0x7540 | SEL);
    return __fmaf_rz(__uint_as_float(bits), 0x1.010102p-8f, -0x1.010102p+15f);
}

// Same for a word that already carries 0x4B in byte 3 and zero in byte 2 (the vertical pass below builds it that way
// for free through the dp2a accumulator): a one-source permute with an immediate selector.
__device__ __forceinline__ float magic_byte1_over_255(uint32_t w) {
    return __fmaf_rz(__uint_as_float(__byte_perm(w, w, 0x3221)), 0x1.010102p-8f, -0x1.010102p+15f);
}

__device__ __forceinline__ void stg_f32(float* base, uint32_t byte_off, float v) {
    // base + byte_off with a 32-bit offset: the compiler forms the address with one carry add per store instead of
    // carrying three 64-bit pointers through the loop.  (An inline-asm st.global here — volatile, so ordered against
    // the volatile shared-memory loads of the next row — cost 23 % on the whole kernel, run r02u.)
    float* p = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + byte_off);
    __builtin_assume(__isGlobal(p));
    *p = v;
}

__device__ __forceinline__ void sts128(uint32_t shared_byte_addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(shared_byte_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t lds32(uint32_t shared_byte_addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_byte_addr));
    return v;
}

// Horizontal 2-tap filter of one source row for this thread's column: three dp2a, pre-shifted by 4
// like OpenCV's vertical pass expects.
__device__ __forceinline__ void lb_hrow(uint32_t p0, uint32_t p1, uint32_t apk, uint32_t (&h)[3]) {
    const uint32_t bg = __byte_perm(p0, p1, 0x5140), rr = __byte_perm(p0, p1, 0x0062);
    h[0] = __dp2a_lo(apk, bg, 0u) >> 4;
    h[1] = __dp2a_hi(apk, bg, 0u) >> 4;
    h[2] = __dp2a_lo(apk, rr, 0u) >> 4;
}

template <bool U8OUT>
__device__ __forceinline__ void lb_store(float* __restrict__ tile_f32, uint8_t* __restrict__ tile_u8, int o, int plane,
                                         int vb, int vg, int vr) {
    if (U8OUT) {
        uint8_t* p = tile_u8 + o * 3;
        p[0] = (uint8_t)vb; p[1] = (uint8_t)vg; p[2] = (uint8_t)vr;
    } else {                                   // planes are R,G,B = source channels 2,1,0
        tile_f32[o] = u8_over_255(vr);
        tile_f32[o + plane] = u8_over_255(vg);
        tile_f32[o + 2 * plane] = u8_over_255(vb);
    }
}

// Phase 1 of the kernel: stage the block's source window into shared memory as pixel words and fill the vertical
// coefficient rows; ends with the CTA-wide barrier.  NT = threads per CTA.
template <int NT>
__device__ __forceinline__ void lb_prologue(const uint8_t* __restrict__ frames, int frame, int64_t frame_bytes, int32_t pitch,
                                            const LbBlock& bd, bool has_src, int smem_row_words, int frames_aligned16,
                                            uint32_t* __restrict__ spix, YCoef* __restrict__ s_y, int oy0, int top, int new_h,
                                            int mode, const YCoef* __restrict__ yt, int sy_lo) {
    // ---- stage source rows [sy_lo, sy_hi], pixels [gx_lo, sx_hi] as pixel words: warp per row, lane per 4-pixel group
    if (has_src) {
        const uint8_t* src = frames + (int64_t)frame * frame_bytes + bd.src_off;
        const int n_groups = bd.n_groups, n_rows = bd.n_rows, row_px = bd.row_px;
        const bool vec_ok = (bd.flags & 2) != 0;
        if ((bd.flags & 4) && frames_aligned16) {
            // 16-pixel groups: three 128-bit loads -> four 128-bit shared stores; up to 3 items per thread in flight.
            // Index arithmetic is kept off the critical instruction count (ncu r02zd: the kernel issues 80 instructions
            // per output pixel at 72 % of the issue slots, 18 of them in this phase): (row, group) of a thread's first
            // item comes from one reciprocal multiply, the next two items advance it by a constant step; the shared
            // address of an item and its bank-swizzle phase travel from the load loop to the store loop packed in
            // ONE register (the compiler otherwise re-derived both from the item index, ~30 instructions per item).
            const int n16 = n_groups >> 2, n_items16 = n_rows * n16;
            const float inv16 = 1.0f / (float)n16;
            const uint32_t spix_sa = (uint32_t)__cvta_generic_to_shared(spix);       // 16-byte aligned: low four bits free
            const uint32_t srow_bytes = 4u * (uint32_t)smem_row_words;               // multiple of 64
            int dr = (int)((float)NT * inv16);                                       // NT items further = dr rows + dg groups
            int dg = NT - dr * n16;
            if (dg < 0) { dr--; dg += n16; }
            if (dg >= n16) { dr++; dg -= n16; }
            for (int it0 = threadIdx.x; it0 < n_items16; it0 += 3 * NT) {
                uint4 q[3][3];
                uint32_t desc[3];                      // shared byte address of the item's 16-word group | swizzle phase (2 bits)
                bool live[3];
                int r = (int)((float)it0 * inv16);
                int g = it0 - r * n16;
                if (g < 0) { r--; g += n16; }
                if (g >= n16) { r++; g -= n16; }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    // (r, g) and the descriptor are made opaque to the optimiser: left alone it re-associated the three
                    // items' index arithmetic into ~75 instructions between the load groups and re-derived it for the stores
                    asm volatile("" : "+r"(r), "+r"(g));
                    live[k] = it0 + k * NT < n_items16;
                    // bank swizzle of the 16-byte chunks inside the 16-word group: chunk ^= (g >> 1) & 3
                    desc[k] = (spix_sa + (uint32_t)r * srow_bytes + 64u * (uint32_t)g) | (((uint32_t)g >> 1) & 3u);
                    asm volatile("" : "+r"(desc[k]));
                    if (live[k]) {
                        const uint4* p128 = reinterpret_cast<const uint4*>(src + (r * pitch + g * 48));
                        q[k][0] = __ldg(p128); q[k][1] = __ldg(p128 + 1); q[k][2] = __ldg(p128 + 2);
                    }
                    g += dg; r += dr;
                    if (g >= n16) { g -= n16; r++; }
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (!live[k]) continue;
                    const uint32_t w[12] = {q[k][0].x, q[k][0].y, q[k][0].z, q[k][0].w, q[k][1].x, q[k][1].y, q[k][1].z, q[k][1].w,
                                            q[k][2].x, q[k][2].y, q[k][2].z, q[k][2].w};
                    const uint32_t sa = desc[k] & ~3u, x = (desc[k] & 3u) << 4;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const uint32_t a = w[3 * t], b = w[3 * t + 1], c = w[3 * t + 2];
                        uint4 o;
                        // The column loop never reads byte 3 of a pixel word, but the masks stay: without them the first
                        // word of a group is a plain register copy that the compiler scheduled straight after its load,
                        // stalling the thread before the loads of its next two items were issued (run r02v: +16 %).
                        o.x = a & 0x00ffffffu;
                        o.y = __byte_perm(a, b, 0x0543) & 0x00ffffffu;
                        o.z = __byte_perm(b, c, 0x0432) & 0x00ffffffu;
                        o.w = c >> 8;
                        sts128(sa + (x ^ (16u * t)), o);
                    }
                }
            }
        } else {
        // flat (row, group) work list, 4 items per thread per pass: all 12 global loads of a pass are
        // issued before the first byte permute consumes one (memory-level parallelism hides DRAM latency)
        const int n_items = n_rows * n_groups;
        const float inv_groups = 1.0f / (float)n_groups;
        for (int it0 = threadIdx.x; it0 < n_items; it0 += 4 * NT) {
            uint32_t a[4], b[4], c[4];
            int soff[4];
            bool fast[4], live[4];
            int poff[4];
            int g4[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int it = it0 + k * NT;
                live[k] = it < n_items;
                int r = (int)((float)it * inv_groups);
                int g = it - r * n_groups;
                if (g < 0) { r--; g += n_groups; }
                if (g >= n_groups) { r++; g -= n_groups; }
                poff[k] = r * pitch + g * 12;
                soff[k] = r * smem_row_words + 4 * (g ^ ((g >> 3) & 3));      // same swizzle as the 16-pixel path
                g4[k] = 4 * g;
                fast[k] = live[k] && vec_ok && 4 * g + 4 <= row_px;
                a[k] = b[k] = c[k] = 0;
                if (fast[k]) {
                    const uint32_t* p32 = reinterpret_cast<const uint32_t*>(src + poff[k]);
                    a[k] = __ldg(p32); b[k] = __ldg(p32 + 1); c[k] = __ldg(p32 + 2);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (!live[k]) continue;
                uint4 w;
                if (fast[k]) {
                    // byte 3 of a pixel word is never read by the column loop, so it is not cleared on this path (the
                    // 32-bit loads land directly in the registers of the 128-bit store: no copy, unlike the path above)
                    w.x = a[k];                                            // B0 G0 R0 (+ B1)
                    w.y = __byte_perm(a[k], b[k], 0x0543);                 // a.b3 b.b0 b.b1 (+ b.b2)
                    w.z = __byte_perm(b[k], c[k], 0x0432);                 // b.b2 b.b3 c.b0 (+ c.b1)
                    w.w = c[k] >> 8;                                       // c.b1 c.b2 c.b3
                } else {                                                   // unaligned tile or the end of a frame row
                    uint32_t v[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        v[q] = 0;
                        if (g4[k] + q < row_px)
                            v[q] = (uint32_t)__ldg(src + poff[k] + 3 * q) | ((uint32_t)__ldg(src + poff[k] + 3 * q + 1) << 8) | ((uint32_t)__ldg(src + poff[k] + 3 * q + 2) << 16);
                    }
                    w = make_uint4(v[0], v[1], v[2], v[3]);
                }
                *reinterpret_cast<uint4*>(spix + soff[k]) = w;
            }
        }
        }
    }
    if (threadIdx.x < kTH) {
        const int dy = oy0 + threadIdx.x - top;
        YCoef yc{0, 0, 0, 0};
        if (has_src && dy >= 0 && dy < new_h) {
            if (mode == MODE_LINEAR) { yc = yt[dy]; yc.r0 -= sy_lo; yc.r1 -= sy_lo; }
            else if (mode == MODE_AREA2) { yc.r0 = 2 * dy - sy_lo; yc.r1 = yc.r0 + 1; }
            else { yc.r0 = dy - sy_lo; yc.r1 = yc.r0; }
            yc.r0 *= 4 * smem_row_words; yc.r1 *= 4 * smem_row_words;      // byte offsets of the two source rows
        }
        s_y[threadIdx.x] = yc;
    }
    __syncthreads();

}

template <bool U8OUT, int MINBLOCKS>
__global__ void __launch_bounds__(kThreads, MINBLOCKS * 256 / kThreads)     // MINBLOCKS counts 256-thread CTAs
letterbox_kernel(const uint8_t* __restrict__ frames, int64_t frame_bytes, int32_t pitch,
                 const LbJob* __restrict__ jobs, const LbBlock* __restrict__ blk2job, int blocks_per_frame,
                 const XCoef* __restrict__ xtab, const YCoef* __restrict__ ytab, int smem_row_words, int frames_aligned16,
                 float* __restrict__ out_f32, uint8_t* __restrict__ out_u8) {
    extern __shared__ __align__(16) uint32_t spix[];      // [rows][smem_row_words] pixel words
    __shared__ YCoef s_y[kTH];                            // vertical coefficients of the block's rows (smem-row relative)

    const int frame = blockIdx.y;                                         // grid = (blocks of a frame, frames): no division
    const LbBlock bd = blk2job[blockIdx.x];                               // two 128-bit loads: everything staging needs
    const uint32_t packed = bd.packed;
    const LbJob& job = jobs[packed >> 24];
    const int oy0 = ((packed >> 12) & 0xfff) * kTH;
    const int ox0 = (packed & 0xfff) * kTW;
    const int mode = job.mode, top = job.top, left = job.left, new_w = job.new_w, new_h = job.new_h;
    const int out_w = job.out_w, out_h = job.out_h, src_w = job.src_w;
    const XCoef* xt = xtab + job.xtab;
    const YCoef* yt = ytab + job.ytab;
    // source window of this block (pixels [gx_lo, sx_hi], rows [sy_lo, sy_hi]); gx_lo is a multiple of 4
    const bool has_src = (bd.flags & 1) != 0;
    const int gx_lo = bd.gx_lo, sy_lo = bd.sy_lo;

    lb_prologue<kThreads>(frames, frame, frame_bytes, pitch, bd, has_src, smem_row_words, frames_aligned16, spix, s_y, oy0, top, new_h,
                          mode, yt, sy_lo);

    // ---- one thread per output column, walking down the block's rows: the horizontal result of a
    // source row is kept in registers and reused when the next output row needs the same row.
    const int ox = ox0 + threadIdx.x;
    if (ox >= out_w) return;
    const int plane = out_h * out_w;
    const int64_t tile_base = job.out_off + (int64_t)frame * job.out_frame_stride;
    float* tile_f32 = U8OUT ? nullptr : out_f32 + tile_base;
    uint8_t* tile_u8 = U8OUT ? out_u8 + tile_base : nullptr;

    const int dx = ox - left;
    const bool cvalid = has_src && dx >= 0 && dx < new_w;
    int i0 = 0, i1 = 0;
    uint32_t apk = 0;
    if (cvalid) {
        if (mode == MODE_LINEAR) {
            const XCoef xc = xt[dx];
            i0 = xc.sx - gx_lo; i1 = min(xc.sx + 1, src_w - 1) - gx_lo; apk = xc.apk;
        } else if (mode == MODE_AREA2) {
            i0 = 2 * dx - gx_lo; i1 = i0 + 1;
        } else {
            i0 = dx - gx_lo;
        }
    }
    const int oy_end = min(oy0 + kTH, out_h);
    int o = oy0 * out_w + ox;

    if (!cvalid) {                                        // padding column (or a block without source pixels)
        for (int oy = oy0; oy < oy_end; oy++, o += out_w) lb_store<U8OUT>(tile_f32, tile_u8, o, plane, 114, 114, 114);
        return;
    }
    // rows of this block that carry image content: [ja, jb); the rest is 114 padding
    const int ja = min(max(top - oy0, 0), oy_end - oy0), jb = max(min(top + new_h - oy0, oy_end - oy0), ja);
    for (int j = 0; j < ja; j++) lb_store<U8OUT>(tile_f32, tile_u8, o + j * out_w, plane, 114, 114, 114);
    for (int j = jb; j < oy_end - oy0; j++) lb_store<U8OUT>(tile_f32, tile_u8, o + j * out_w, plane, 114, 114, 114);
    o += ja * out_w;

    // raw shared-memory byte addresses of the two taps of this column (row offsets are added per row)
    // staged rows are stored with their 16-byte chunks XOR-swizzled inside each 16-word group (chunk ^= (word >> 5) & 3)
    // so that the 64-byte-strided 128-bit staging stores of a quarter-warp fall into distinct banks
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(spix) + 4u * ((uint32_t)i0 ^ ((((uint32_t)i0 >> 5) & 3u) << 2));
    const uint32_t a1 = (uint32_t)__cvta_generic_to_shared(spix) + 4u * ((uint32_t)i1 ^ ((((uint32_t)i1 >> 5) & 3u) << 2));
    // Output addressing: three plane bases (R, G, B = source channels 2, 1, 0) and ONE 32-bit byte offset that
    // advances by a row per iteration; each store address costs one IADD3 + one IMAD.X (see stg_f32).
    float* const q0 = U8OUT ? nullptr : tile_f32;
    float* const q1 = U8OUT ? nullptr : q0 + plane;
    float* const q2 = U8OUT ? nullptr : q1 + plane;
    uint32_t ob = 4u * (uint32_t)o;
    const uint32_t ob_step = 4u * (uint32_t)out_w;
    auto next_row = [&]() { ob += ob_step; };
    uint8_t* q8 = U8OUT ? tile_u8 + 3 * o : nullptr;

    // one output pixel whose channel values are bytes SB / SG / SR of the words wb / wg / wr
    auto emit = [&](uint32_t wb, auto sb, uint32_t wg, auto sg, uint32_t wr, auto sr) {
        constexpr int SB = decltype(sb)::value, SG = decltype(sg)::value, SR = decltype(sr)::value;
        if (U8OUT) {
            q8[0] = (uint8_t)(wb >> (8 * SB)); q8[1] = (uint8_t)(wg >> (8 * SG)); q8[2] = (uint8_t)(wr >> (8 * SR));
            q8 += 3 * out_w;
        } else {
            stg_f32(q0, ob, byte_over_255<SR>(wr));
            stg_f32(q1, ob, byte_over_255<SG>(wg));
            stg_f32(q2, ob, byte_over_255<SB>(wb));
            next_row();
        }
    };
    using B0 = std::integral_constant<int, 0>;
    using B1 = std::integral_constant<int, 1>;
    using B2 = std::integral_constant<int, 2>;

    if (mode == MODE_COPY) {
#pragma unroll kRowUnroll
        for (int j = ja; j < jb; j++) {
            const uint32_t p = lds32(a0 + (uint32_t)s_y[j].r0);
            emit(p, B0{}, p, B1{}, p, B2{});
        }
    } else if (mode == MODE_AREA2) {
#pragma unroll kRowUnroll
        for (int j = ja; j < jb; j++) {
            const YCoef yc = s_y[j];
            const uint32_t p00 = lds32(a0 + yc.r0), p01 = lds32(a1 + yc.r0), p10 = lds32(a0 + yc.r1), p11 = lds32(a1 + yc.r1);
            // B and R ride in the two 16-bit lanes of one word, G in another; (sum + 2) >> 2 lands in bytes 0 / 2
            // after the shift (lane sums <= 1022: ten bits, the upper eight are the value)
            const uint32_t br = (p00 & 0x00ff00ffu) + (p01 & 0x00ff00ffu) + (p10 & 0x00ff00ffu) + (p11 & 0x00ff00ffu) + 0x00020002u;
            const uint32_t gg = ((p00 >> 8) & 255) + ((p01 >> 8) & 255) + ((p10 >> 8) & 255) + ((p11 >> 8) & 255) + 2;
            emit(br >> 2, B0{}, gg >> 2, B0{}, br >> 2, B2{});
        }
    } else {
        uint32_t ph[3] = {0u, 0u, 0u};
        int prev_r1 = -1;
#pragma unroll kRowUnroll
        for (int j = ja; j < jb; j++) {
            const YCoef yc = s_y[j];
            uint32_t h0[3], h1[3];
            if (yc.r0 == prev_r1) {                       // warp-uniform: the upper source row is the previous row's lower one
                h0[0] = ph[0]; h0[1] = ph[1]; h0[2] = ph[2];
            } else {
                lb_hrow(lds32(a0 + yc.r0), lds32(a1 + yc.r0), apk, h0);
            }
            lb_hrow(lds32(a0 + yc.r1), lds32(a1 + yc.r1), apk, h1);
            ph[0] = h1[0]; ph[1] = h1[1]; ph[2] = h1[2];
            prev_r1 = yc.r1;
            const uint32_t c0 = (uint32_t)yc.b0, c1 = (uint32_t)yc.b1;
            // OpenCV vertical pass: (((c0*(H0>>4))>>16) + ((c1*(H1>>4))>>16) + 2) >> 2.  The "+2" rides on the first
            // product as 2<<16 (no carry into it: products < 2^27), one byte permute pairs the two upper halves, and a
            // dp2a by (64, 64) adds them AND moves the final ">>2" to a byte boundary: the pixel value is byte 1 of the
            // result (64 * sum <= 64 * 1022 < 2^16; the accumulator plants the 2^23 exponent byte for the float path).
            // Per channel: 2 IMAD + 1 IDP on the FMA pipe and 1 PRMT on the ALU pipe, instead of SHF, LEA.HI, VIADD, SHF on
            // the ALU pipe (ncu on the round-1 loop: ALU pipe 71 % busy; run r02w: -3 % alone, 0.74 -> 0.79 in the step).
            uint32_t v[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint32_t hi2 = __byte_perm(c0 * h0[c] + 0x20000u, c1 * h1[c], 0x7632);
                v[c] = __dp2a_lo(hi2, 0x4040u, U8OUT ? 0u : 0x4B000000u);
            }
            if (U8OUT) {
                emit(v[0], B1{}, v[1], B1{}, v[2], B1{});
            } else {                                      // byte 3 = 0x4B, byte 2 = 0, byte 1 = the value
                stg_f32(q0, ob, magic_byte1_over_255(v[2]));
                stg_f32(q1, ob, magic_byte1_over_255(v[1]));
                stg_f32(q2, ob, magic_byte1_over_255(v[0]));
                next_row();
            }
        }
    }
}

// ------------------------------------------------------------------------------------ host side
struct Geometry {           // ultralytics LetterBox.__call__ for one source size
    int new_w, new_h, top, left, out_h, out_w;
    float gain, pad_x, pad_y;
};

// Python's round() (banker's) on a double == nearbyint() in the default rounding mode.
inline int py_round(double x) { return (int)nearbyint(x); }

Geometry letterbox_geometry(int h, int w, int imgsz, bool auto_pad, int stride) {
    Geometry g;
    double r = std::min((double)imgsz / h, (double)imgsz / w);
    g.new_w = py_round(w * r);
    g.new_h = py_round(h * r);
    int idw = imgsz - g.new_w, idh = imgsz - g.new_h;
    if (auto_pad) { idw = ((idw % stride) + stride) % stride; idh = ((idh % stride) + stride) % stride; }
    double dw = idw / 2.0, dh = idh / 2.0;
    g.top = py_round(dh - 0.1);
    int bottom = py_round(dh + 0.1);
    g.left = py_round(dw - 0.1);
    int right = py_round(dw + 0.1);
    g.out_h = g.new_h + g.top + bottom;
    g.out_w = g.new_w + g.left + right;
    // ultralytics scale_boxes(img1_shape=(out_h,out_w), boxes, img0_shape=(h,w))
    double gain = std::min((double)g.out_h / h, (double)g.out_w / w);
    g.gain = (float)gain;
    g.pad_x = (float)py_round((g.out_w - w * gain) / 2 - 0.1);
    g.pad_y = (float)py_round((g.out_h - h * gain) / 2 - 0.1);
    return g;
}

// OpenCV resizeGeneric_ coefficient set-up for INTER_LINEAR (float32 maths, 11-bit fixed point).
void linear_coeffs(int src, int dst, bool clamp_fraction, std::vector<int>& idx, std::vector<int>& c0,
                   std::vector<int>& c1) {
    idx.resize(dst); c0.resize(dst); c1.resize(dst);
    double scale = (double)src / dst;
    for (int d = 0; d < dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (clamp_fraction) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= src - 1) { s = src - 1; f = 0.f; }
        }
        idx[d] = s;
        c0[d] = (int)nearbyintf((1.f - f) * 2048.f);
        c1[d] = (int)nearbyintf(f * 2048.f);
    }
}

}  // namespace

struct hvb_lb_plan {
    hvb_ctx* ctx = nullptr;
    int n_frames = 0, frame_h = 0, frame_w = 0, mode = 0;
    std::vector<hvb_lb_class> classes;
    std::vector<hvb_lb_tile> tiles;       // n_frames * tiles_per_frame, frame-major
    int tiles_per_frame = 0;
    int blocks_per_frame = 0;
    int smem_row_stride = 0, smem_bytes = 0;
    int64_t out_elems = 0, read_bytes = 0, write_bytes = 0;
    void* dev = nullptr;                  // one allocation: jobs | blk2job | xtab | ytab
    LbJob* jobs_dev = nullptr;
    LbBlock* blk2job_dev = nullptr;
    XCoef* xtab_dev = nullptr;
    YCoef* ytab_dev = nullptr;
};

extern "C" {

int hvb_lb_plan_create(hvb_ctx* ctx, int n_frames, int frame_h, int frame_w, int mode, int imgsz, int auto_pad,
                       int stride, int slice_w, int slice_h, int overlap_w, int overlap_h, hvb_lb_plan** out_plan) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(out_plan != nullptr, "null out_plan");
    *out_plan = nullptr;
    HVB_ARG(n_frames > 0 && frame_h > 0 && frame_w > 0, "bad frame geometry");
    HVB_ARG(imgsz > 0 && stride > 0, "bad imgsz/stride");
    HVB_ARG(mode >= HVB_LB_WHOLE && mode <= HVB_LB_SLICE_UNIFORM, "bad mode");
    if (mode != HVB_LB_WHOLE) {
        HVB_ARG(slice_w > 0 && slice_h > 0 && overlap_w >= 0 && overlap_h >= 0 && overlap_w < slice_w &&
                    overlap_h < slice_h, "bad slice geometry");
    }
    hvb_lb_plan* p = new hvb_lb_plan();
    p->ctx = ctx; p->n_frames = n_frames; p->frame_h = frame_h; p->frame_w = frame_w; p->mode = mode;

    // ---- tiles of one frame (supervision InferenceSlicer._generate_offset: clipped, row-major y then x)
    struct Src { int x, y, w, h; };
    std::vector<Src> srcs;
    if (mode == HVB_LB_WHOLE) {
        srcs.push_back({0, 0, frame_w, frame_h});
    } else {
        for (int y = 0; y < frame_h; y += slice_h - overlap_h)
            for (int x = 0; x < frame_w; x += slice_w - overlap_w)
                srcs.push_back({x, y, std::min(x + slice_w, frame_w) - x, std::min(y + slice_h, frame_h) - y});
    }
    const int T = (int)srcs.size();
    p->tiles_per_frame = T;
    if (T > 255) { delete p; hvb_set_error("more than 255 tiles per frame"); return HVB_ERR_CAPACITY; }

    const bool use_auto = (mode == HVB_LB_SLICE_UNIFORM) ? false : (auto_pad != 0);
    std::vector<Geometry> geo(T);
    std::vector<int> cls_of(T), k_of(T);
    for (int t = 0; t < T; t++) {
        geo[t] = letterbox_geometry(srcs[t].h, srcs[t].w, imgsz, use_auto, stride);
        int c = -1;
        for (size_t i = 0; i < p->classes.size(); i++)
            if (p->classes[i].out_h == geo[t].out_h && p->classes[i].out_w == geo[t].out_w) c = (int)i;
        if (c < 0) {
            hvb_lb_class k{};
            k.out_h = geo[t].out_h; k.out_w = geo[t].out_w; k.tiles_per_frame = 0;
            p->classes.push_back(k);
            c = (int)p->classes.size() - 1;
        }
        cls_of[t] = c;
        k_of[t] = p->classes[c].tiles_per_frame++;
    }
    int64_t off = 0;
    for (auto& c : p->classes) {
        c.batch = c.tiles_per_frame * n_frames;
        c.out_offset = off;
        off += (int64_t)c.batch * 3 * c.out_h * c.out_w;
    }
    p->out_elems = off;

    // ---- jobs, coefficient tables (shared between identical geometries), block work list
    std::vector<LbJob> jobs(T);
    std::vector<XCoef> xt;
    std::vector<YCoef> yt;
    std::map<std::pair<int, int>, int> xkey, ykey;    // (src,dst) -> table offset
    std::vector<LbBlock> blk;
    int max_rows = 1, max_span = 4;
    for (int t = 0; t < T; t++) {
        const Geometry& g = geo[t];
        LbJob& j = jobs[t];
        j.src_off = ((int64_t)srcs[t].y * frame_w + srcs[t].x) * 3;
        const hvb_lb_class& c = p->classes[cls_of[t]];
        const int64_t tile_elems = (int64_t)3 * c.out_h * c.out_w;
        j.out_off = c.out_offset + (int64_t)k_of[t] * tile_elems;
        j.out_frame_stride = (int64_t)c.tiles_per_frame * tile_elems;
        j.src_w = srcs[t].w; j.src_h = srcs[t].h; j.new_w = g.new_w; j.new_h = g.new_h;
        j.top = g.top; j.left = g.left; j.out_h = g.out_h; j.out_w = g.out_w;
        j.xtab = j.ytab = 0; j.pad_[0] = j.pad_[1] = j.pad_[2] = 0;
        j.row_px = frame_w - srcs[t].x;
        j.vec_ok = ((frame_w * 3) % 4 == 0 && j.src_off % 4 == 0 && ((int64_t)frame_h * frame_w * 3) % 4 == 0) ? 1 : 0;
        if (g.new_w == srcs[t].w && g.new_h == srcs[t].h) j.mode = MODE_COPY;
        else if (srcs[t].w == 2 * g.new_w && srcs[t].h == 2 * g.new_h) j.mode = MODE_AREA2;
        else {
            j.mode = MODE_LINEAR;
            auto kx = std::make_pair(srcs[t].w, g.new_w);
            if (!xkey.count(kx)) {
                xkey[kx] = (int)xt.size();
                std::vector<int> s, a0, a1;
                linear_coeffs(srcs[t].w, g.new_w, true, s, a0, a1);
                for (int d = 0; d < g.new_w; d++) xt.push_back({s[d], (uint32_t)a0[d] | ((uint32_t)a1[d] << 16)});
            }
            auto ky = std::make_pair(srcs[t].h, g.new_h);
            if (!ykey.count(ky)) {
                ykey[ky] = (int)yt.size();
                std::vector<int> s, b0, b1;
                linear_coeffs(srcs[t].h, g.new_h, false, s, b0, b1);
                for (int d = 0; d < g.new_h; d++) {
                    int r0 = std::min(std::max(s[d], 0), srcs[t].h - 1);
                    int r1 = std::min(std::max(s[d] + 1, 0), srcs[t].h - 1);
                    yt.push_back({r0, r1, b0[d], b1[d]});
                }
            }
            j.xtab = xkey[kx]; j.ytab = ykey[ky];
        }
        // block list of this job with each block's source window; shared-memory need = the worst block
        const int by = hvb_div_up(g.out_h, kTH), bx = hvb_div_up(g.out_w, kTW);
        if (by > 4095 || bx > 4095) { delete p; hvb_set_error("output too large for the block table"); return HVB_ERR_CAPACITY; }
        for (int y = 0; y < by; y++)
            for (int x = 0; x < bx; x++) {
                LbBlock b{};
                b.packed = ((uint32_t)t << 24) | ((uint32_t)y << 12) | (uint32_t)x;
                const int oy0 = y * kTH, ox0 = x * kTW;
                const int dy_lo = std::max(oy0 - g.top, 0), dy_hi = std::min(oy0 + kTH - g.top, g.new_h);
                const int dx_lo = std::max(ox0 - g.left, 0), dx_hi = std::min(ox0 + kTW - g.left, g.new_w);
                if (dy_lo < dy_hi && dx_lo < dx_hi) {
                    int sx_hi, sy_hi;
                    if (j.mode == MODE_LINEAR) {
                        b.gx_lo = xt[j.xtab + dx_lo].sx;
                        sx_hi = std::min(xt[j.xtab + dx_hi - 1].sx + 1, srcs[t].w - 1);
                        b.sy_lo = yt[j.ytab + dy_lo].r0;
                        sy_hi = yt[j.ytab + dy_hi - 1].r1;
                    } else if (j.mode == MODE_AREA2) {
                        b.gx_lo = 2 * dx_lo; sx_hi = 2 * dx_hi - 1; b.sy_lo = 2 * dy_lo; sy_hi = 2 * dy_hi - 1;
                    } else {
                        b.gx_lo = dx_lo; sx_hi = dx_hi - 1; b.sy_lo = dy_lo; sy_hi = dy_hi - 1;
                    }
                    b.gx_lo &= ~3;
                    b.flags = 1 | (j.vec_ok ? 2 : 0);
                    // 16-pixel (48-byte) groups when every row start of the window is 16-byte aligned and the
                    // widened window stays inside the frame row
                    {
                        const int g16 = b.gx_lo & ~15;
                        const int n16 = ((sx_hi - g16) >> 4) + 1;
                        const int64_t off16 = j.src_off + (int64_t)g16 * 3;
                        if ((frame_w * 3) % 16 == 0 && off16 % 16 == 0 && ((int64_t)frame_h * frame_w * 3) % 16 == 0 &&
                            g16 + 16 * n16 <= j.row_px) {
                            b.gx_lo = g16;
                            b.flags |= 4;
                            sx_hi = g16 + 16 * n16 - 1;
                        }
                    }
                    b.n_rows = sy_hi - b.sy_lo + 1;
                    b.n_groups = ((sx_hi - b.gx_lo) >> 2) + 1;
                    b.row_px = j.row_px - b.gx_lo;
                    b.src_off = j.src_off + (int64_t)b.sy_lo * frame_w * 3 + (int64_t)b.gx_lo * 3;
                    max_rows = std::max(max_rows, b.n_rows);
                    max_span = std::max(max_span, 4 * b.n_groups);
                }
                blk.push_back(b);
            }

        p->read_bytes += (int64_t)srcs[t].w * srcs[t].h * 3;
        p->write_bytes += (int64_t)3 * g.out_h * g.out_w * 4;
    }
    // every source byte counted once per frame (tiles overlap): algorithmic read = the frame itself
    p->read_bytes = std::min<int64_t>(p->read_bytes, (int64_t)frame_h * frame_w * 3) * n_frames;
    p->write_bytes *= n_frames;
    p->blocks_per_frame = (int)blk.size();
    p->smem_row_stride = (max_span + 15) & ~15;                        // pixel words per staged row: whole 16-word groups (the bank swizzle permutes chunks inside a group)
    p->smem_bytes = p->smem_row_stride * max_rows * 4;
    if (p->smem_bytes > ctx->max_smem_optin - 1024) {
        delete p;
        hvb_set_error("letterbox block needs %d bytes of shared memory (down-scale factor too large)", p->smem_bytes);
        return HVB_ERR_CAPACITY;
    }
    if (n_frames > 65535) {           // the launch grid is (blocks of a frame, frames): gridDim.y
        delete p; hvb_set_error("more than 65535 frames per letterbox plan: split the chunk"); return HVB_ERR_CAPACITY;
    }

    // ---- host-visible tile list (frame-major) for the decode stage
    p->tiles.resize((size_t)T * n_frames);
    for (int f = 0; f < n_frames; f++)
        for (int t = 0; t < T; t++) {
            hvb_lb_tile& o = p->tiles[(size_t)f * T + t];
            const Geometry& g = geo[t];
            o.frame = f; o.tile = t; o.cls = cls_of[t];
            o.batch_index = f * p->classes[cls_of[t]].tiles_per_frame + k_of[t];
            o.src_x = srcs[t].x; o.src_y = srcs[t].y; o.src_w = srcs[t].w; o.src_h = srcs[t].h;
            o.new_w = g.new_w; o.new_h = g.new_h; o.top = g.top; o.left = g.left;
            o.out_h = g.out_h; o.out_w = g.out_w; o.gain = g.gain; o.pad_x = g.pad_x; o.pad_y = g.pad_y;
        }

    // ---- upload
    if (xt.empty()) xt.push_back({0, 0u});
    if (yt.empty()) yt.push_back({0, 0, 0, 0});
    size_t o_jobs = 0;
    size_t o_blk = o_jobs + ((jobs.size() * sizeof(LbJob) + 255) & ~(size_t)255);
    size_t o_x = o_blk + ((blk.size() * sizeof(LbBlock) + 255) & ~(size_t)255);
    size_t o_y = o_x + ((xt.size() * sizeof(XCoef) + 255) & ~(size_t)255);
    size_t total = o_y + yt.size() * sizeof(YCoef);
    std::vector<uint8_t> host(total, 0);
    memcpy(host.data() + o_jobs, jobs.data(), jobs.size() * sizeof(LbJob));
    memcpy(host.data() + o_blk, blk.data(), blk.size() * sizeof(LbBlock));
    memcpy(host.data() + o_x, xt.data(), xt.size() * sizeof(XCoef));
    memcpy(host.data() + o_y, yt.data(), yt.size() * sizeof(YCoef));
    cudaError_t e = cudaMalloc(&p->dev, total);
    if (e != cudaSuccess) { delete p; return hvb_cuda_fail(e, "cudaMalloc(plan)", __FILE__, __LINE__); }
    e = cudaMemcpy(p->dev, host.data(), total, cudaMemcpyHostToDevice);
    // a pageable-source cudaMemcpy may return before the DMA lands, and the kernels run on non-blocking streams that do
    // not order against the legacy stream: finish the upload here (plan creation is a one-off)
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(p->dev); delete p; return hvb_cuda_fail(e, "cudaMemcpy(plan)", __FILE__, __LINE__); }
    p->jobs_dev = (LbJob*)((uint8_t*)p->dev + o_jobs);
    p->blk2job_dev = (LbBlock*)((uint8_t*)p->dev + o_blk);
    p->xtab_dev = (XCoef*)((uint8_t*)p->dev + o_x);
    p->ytab_dev = (YCoef*)((uint8_t*)p->dev + o_y);
    if (p->smem_bytes > 48 * 1024) {
        HVB_CUDA(cudaFuncSetAttribute(letterbox_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes));
        HVB_CUDA(cudaFuncSetAttribute(letterbox_kernel<true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes));
    }
    *out_plan = p;
    return HVB_OK;
}

int hvb_lb_plan_destroy(hvb_lb_plan* plan) {
    if (!plan) return HVB_OK;
    cudaSetDevice(plan->ctx->device);
    if (plan->dev) cudaFree(plan->dev);
    delete plan;
    return HVB_OK;
}

int hvb_lb_plan_num_classes(const hvb_lb_plan* plan, int* out_n) {
    if (!plan || !out_n) { hvb_set_error("null argument"); return HVB_ERR_ARG; }
    *out_n = (int)plan->classes.size();
    return HVB_OK;
}

int hvb_lb_plan_get_class(const hvb_lb_plan* plan, int cls, hvb_lb_class* out_class) {
    if (!plan || !out_class) { hvb_set_error("null argument"); return HVB_ERR_ARG; }
    if (cls < 0 || cls >= (int)plan->classes.size()) { hvb_set_error("class index out of range"); return HVB_ERR_ARG; }
    *out_class = plan->classes[cls];
    return HVB_OK;
}

int hvb_lb_plan_num_tiles(const hvb_lb_plan* plan, int* out_n) {
    if (!plan || !out_n) { hvb_set_error("null argument"); return HVB_ERR_ARG; }
    *out_n = (int)plan->tiles.size();
    return HVB_OK;
}

int hvb_lb_plan_get_tiles(const hvb_lb_plan* plan, hvb_lb_tile* out_tiles_host, int capacity) {
    if (!plan || !out_tiles_host) { hvb_set_error("null argument"); return HVB_ERR_ARG; }
    if (capacity < (int)plan->tiles.size()) { hvb_set_error("tile buffer too small"); return HVB_ERR_ARG; }
    memcpy(out_tiles_host, plan->tiles.data(), plan->tiles.size() * sizeof(hvb_lb_tile));
    return HVB_OK;
}

int hvb_lb_plan_out_floats(const hvb_lb_plan* plan, int64_t* out_floats) {
    if (!plan || !out_floats) { hvb_set_error("null argument"); return HVB_ERR_ARG; }
    *out_floats = plan->out_elems;
    return HVB_OK;
}

int hvb_lb_plan_bytes(const hvb_lb_plan* plan, int64_t* out_read_bytes, int64_t* out_write_bytes) {
    if (!plan) { hvb_set_error("null argument"); return HVB_ERR_ARG; }
    if (out_read_bytes) *out_read_bytes = plan->read_bytes;
    if (out_write_bytes) *out_write_bytes = plan->write_bytes;
    return HVB_OK;
}

static int lb_run(hvb_lb_plan* p, const uint8_t* frames_dev, float* out_f32, uint8_t* out_u8) {
    if (!p) { hvb_set_error("null plan"); return HVB_ERR_ARG; }
    hvb_ctx* ctx = p->ctx;
    HVB_CHECK_CTX(ctx);
    HVB_ARG(frames_dev && (out_f32 || out_u8), "null buffer");
    HVB_ARG(((uintptr_t)frames_dev & 3) == 0, "frames_dev must be 4-byte aligned");
    HVB_ARG(p->n_frames <= 65535, "more than 65535 frames per launch");
    const dim3 grid((unsigned)p->blocks_per_frame, (unsigned)p->n_frames);
    const int64_t frame_bytes = (int64_t)p->frame_h * p->frame_w * 3;
    const int al16 = (((uintptr_t)frames_dev & 15) == 0) ? 1 : 0;
    // 64 registers / eight 128-thread CTAs per SM for every float32 plan.  (With 256-thread blocks the copy-dominated slice
    // plans ran best at 48 registers / 5 CTAs per SM; with 128-thread blocks 64 registers win for both kinds — run r02y:
    // K1b 479 -> 463 us per 16 4K frames, K1a 186 against 221 us with the 48-register build.)
    if (out_u8)
        letterbox_kernel<true, 5><<<grid, kThreads, p->smem_bytes, ctx->stream>>>(
            frames_dev, frame_bytes, p->frame_w * 3, p->jobs_dev, p->blk2job_dev, p->blocks_per_frame, p->xtab_dev,
            p->ytab_dev, p->smem_row_stride, al16, nullptr, out_u8);
    else
        letterbox_kernel<false, 4><<<grid, kThreads, p->smem_bytes, ctx->stream>>>(
            frames_dev, frame_bytes, p->frame_w * 3, p->jobs_dev, p->blk2job_dev, p->blocks_per_frame, p->xtab_dev,
            p->ytab_dev, p->smem_row_stride, al16, out_f32, nullptr);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_lb_plan_run(hvb_lb_plan* plan, const uint8_t* frames_dev, float* out_dev) {
    return lb_run(plan, frames_dev, out_dev, nullptr);
}

int hvb_lb_plan_run_u8(hvb_lb_plan* plan, const uint8_t* frames_dev, uint8_t* out_dev) {
    return lb_run(plan, frames_dev, nullptr, out_dev);
}

}  // extern "C"

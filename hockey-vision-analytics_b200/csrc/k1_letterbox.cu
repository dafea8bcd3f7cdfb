// K1a / K1b — BGR uint8 frames -> letterboxed, RGB, /255, NCHW float32 batches, for whole frames
// (ultralytics LetterBox + preprocess behind hockey/main.py:179-184) and for InferenceSlicer
// tiles (supervision _generate_offset + crop_image, then LetterBox per tile; README.md:25).
//
// The uint8 stage is bit-exact with cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder(114)
// (SURVEY.md App. A5): 11-bit fixed-point coefficients built in float32 exactly like OpenCV
// (on the host, once per geometry), horizontal fraction clamped at the borders, vertical indices
// clamped with the fraction kept, the (>>4, >>16, +2 >>2) vertical rounding, and the 2x2
// area-average shortcut when both scale factors are exactly 2.  The float stage reproduces
// numpy/torch `x / 255` in float32 exactly with a reciprocal multiply plus two FMA corrections
// (verified exhaustively for the 256 possible inputs).
//
// One launch covers a chunk of frames: grid = blocks_per_frame x n_frames.  Each CTA produces an
// 8-row x 256-column block of one tile's output: it first stages the source bytes the block needs
// into shared memory with coalesced 128-bit loads (the source rows are shared by the 3 output
// planes, by neighbouring output pixels and by the two taps of neighbouring output rows), then
// every thread produces 4 consecutive pixels x 3 planes and writes them with 128-bit stores.
#include "hvb_common.cuh"

#include <algorithm>
#include <math.h>
#include <map>

namespace {

constexpr int kTH = 8;          // output rows per CTA
constexpr int kTW = 256;        // output columns per CTA
constexpr int kThreads = 256;   // 64 column groups of 4 px  x  4 row groups (2 rows each)

enum { MODE_COPY = 0, MODE_LINEAR = 1, MODE_AREA2 = 2 };

struct LbJob {
    int64_t src_off;     // byte offset of the tile's first source pixel inside frame 0
    int64_t out_off;     // element offset of this tile's output block for frame 0
    int64_t out_frame_stride;  // elements between the same tile of consecutive frames
    int32_t src_w, src_h, new_w, new_h, top, left, out_h, out_w;
    int32_t mode;
    int32_t xtab, ytab;  // offsets (entries) of this job's coefficient tables
    int32_t pad_;
};

struct XCoef { int32_t sx; int16_t a0, a1; };            // 8 bytes
struct YCoef { int32_t r0, r1; int32_t b0, b1; };        // 16 bytes

__device__ __forceinline__ float u8_over_255(int v) {
    // float32(v) / 255.0f, correctly rounded: q = v*rcp; r = v - q*255; q += r*rcp
    const float rcp = 1.0f / 255.0f;
    float x = __int_as_float(0x4B000000 | v) - 8388608.0f;
    float q = __fmul_rn(x, rcp);
    float r = __fmaf_rn(-q, 255.0f, x);
    return __fmaf_rn(r, rcp, q);
}

template <bool U8OUT>
__global__ void __launch_bounds__(kThreads)
letterbox_kernel(const uint8_t* __restrict__ frames, int64_t frame_bytes, int32_t pitch,
                 const LbJob* __restrict__ jobs, const uint32_t* __restrict__ blk2job, int blocks_per_frame,
                 const XCoef* __restrict__ xtab, const YCoef* __restrict__ ytab, int smem_row_stride,
                 float* __restrict__ out_f32, uint8_t* __restrict__ out_u8) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_geo[8];   // sx_lo, sx_hi, sy_lo, sy_hi

    const int frame = blockIdx.x / blocks_per_frame;
    const uint32_t packed = __ldg(blk2job + (blockIdx.x - frame * blocks_per_frame));
    const LbJob job = jobs[packed >> 24];
    const int oy0 = ((packed >> 12) & 0xfff) * kTH;
    const int ox0 = (packed & 0xfff) * kTW;

    // valid (non-padding) output window of this block, in resized-image coordinates
    const int dy_lo = max(oy0 - job.top, 0), dy_hi = min(oy0 + kTH - job.top, job.new_h);     // [lo,hi)
    const int dx_lo = max(ox0 - job.left, 0), dx_hi = min(ox0 + kTW - job.left, job.new_w);
    const bool has_src = dy_lo < dy_hi && dx_lo < dx_hi;

    if (threadIdx.x == 0 && has_src) {
        int sx_lo, sx_hi, sy_lo, sy_hi;
        if (job.mode == MODE_LINEAR) {
            sx_lo = xtab[job.xtab + dx_lo].sx;
            sx_hi = min(xtab[job.xtab + dx_hi - 1].sx + 1, job.src_w - 1);
            sy_lo = ytab[job.ytab + dy_lo].r0;
            sy_hi = ytab[job.ytab + dy_hi - 1].r1;
        } else if (job.mode == MODE_AREA2) {
            sx_lo = 2 * dx_lo; sx_hi = 2 * dx_hi - 1; sy_lo = 2 * dy_lo; sy_hi = 2 * dy_hi - 1;
        } else {
            sx_lo = dx_lo; sx_hi = dx_hi - 1; sy_lo = dy_lo; sy_hi = dy_hi - 1;
        }
        s_geo[0] = sx_lo; s_geo[1] = sx_hi; s_geo[2] = sy_lo; s_geo[3] = sy_hi;
    }
    __syncthreads();
    const int sx_lo = s_geo[0], sx_hi = s_geo[1], sy_lo = s_geo[2], sy_hi = s_geo[3];
    const uint8_t* src = frames + (int64_t)frame * frame_bytes + job.src_off;

    // ---- stage source rows [sy_lo, sy_hi], bytes [sx_lo*3, (sx_hi+1)*3) with aligned 16-byte loads
    if (has_src) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int span = (sx_hi - sx_lo + 1) * 3;
        for (int r = sy_lo + warp; r <= sy_hi; r += kThreads / 32) {
            const uint8_t* g = src + (int64_t)r * pitch + sx_lo * 3;
            const int shift = (int)((uintptr_t)g & 15);
            const uint4* g16 = reinterpret_cast<const uint4*>(g - shift);
            uint4* s16 = reinterpret_cast<uint4*>(smem + (r - sy_lo) * smem_row_stride);
            const int nchunk = (shift + span + 15) >> 4;
            for (int c = lane; c < nchunk; c += 32) s16[c] = __ldg(g16 + c);
        }
    }
    __syncthreads();

    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int ox = ox0 + tx * 4;
    if (ox >= job.out_w) return;
    const int64_t plane = (int64_t)job.out_h * job.out_w;
    const int64_t tile_base = job.out_off + (int64_t)frame * job.out_frame_stride;

    // per-thread column coefficients (4 pixels)
    int csx[4], ca0[4], ca1[4];
    bool cvalid[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int dx = ox + i - job.left;
        cvalid[i] = dx >= 0 && dx < job.new_w && (ox + i) < job.out_w;
        csx[i] = 0; ca0[i] = 0; ca1[i] = 0;
        if (cvalid[i]) {
            if (job.mode == MODE_LINEAR) {
                XCoef xc = xtab[job.xtab + dx];
                csx[i] = xc.sx; ca0[i] = xc.a0; ca1[i] = xc.a1;
            } else if (job.mode == MODE_AREA2) {
                csx[i] = 2 * dx;
            } else {
                csx[i] = dx;
            }
        }
    }

#pragma unroll
    for (int j = 0; j < kTH / 4; j++) {
        const int oy = oy0 + ty + 4 * j;
        if (oy >= job.out_h) break;
        const int dy = oy - job.top;
        const bool rvalid = dy >= 0 && dy < job.new_h;
        int val[4][3];
#pragma unroll
        for (int i = 0; i < 4; i++) { val[i][0] = val[i][1] = val[i][2] = 114; }

        if (rvalid) {
            if (job.mode == MODE_COPY) {
                const uint8_t* row = smem + (dy - sy_lo) * smem_row_stride +
                                     (int)((uintptr_t)(src + (int64_t)dy * pitch + sx_lo * 3) & 15);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (cvalid[i]) {
                        const uint8_t* p = row + (csx[i] - sx_lo) * 3;
                        val[i][0] = p[0]; val[i][1] = p[1]; val[i][2] = p[2];
                    }
            } else if (job.mode == MODE_AREA2) {
                const int r0 = 2 * dy;
                const uint8_t* row0 = smem + (r0 - sy_lo) * smem_row_stride +
                                      (int)((uintptr_t)(src + (int64_t)r0 * pitch + sx_lo * 3) & 15);
                const uint8_t* row1 = smem + (r0 + 1 - sy_lo) * smem_row_stride +
                                      (int)((uintptr_t)(src + (int64_t)(r0 + 1) * pitch + sx_lo * 3) & 15);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (cvalid[i]) {
                        const int o = (csx[i] - sx_lo) * 3;
#pragma unroll
                        for (int c = 0; c < 3; c++)
                            val[i][c] = (row0[o + c] + row0[o + 3 + c] + row1[o + c] + row1[o + 3 + c] + 2) >> 2;
                    }
            } else {
                const YCoef yc = ytab[job.ytab + dy];
                const uint8_t* row0 = smem + (yc.r0 - sy_lo) * smem_row_stride +
                                      (int)((uintptr_t)(src + (int64_t)yc.r0 * pitch + sx_lo * 3) & 15);
                const uint8_t* row1 = smem + (yc.r1 - sy_lo) * smem_row_stride +
                                      (int)((uintptr_t)(src + (int64_t)yc.r1 * pitch + sx_lo * 3) & 15);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (cvalid[i]) {
                        const int o0 = (csx[i] - sx_lo) * 3;
                        const int o1 = (min(csx[i] + 1, job.src_w - 1) - sx_lo) * 3;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            int h0 = row0[o0 + c] * ca0[i] + row0[o1 + c] * ca1[i];
                            int h1 = row1[o0 + c] * ca0[i] + row1[o1 + c] * ca1[i];
                            val[i][c] = (((yc.b0 * (h0 >> 4)) >> 16) + ((yc.b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                        }
                    }
            }
        }

        if (U8OUT) {
            uint8_t* o = out_u8 + tile_base + ((int64_t)oy * job.out_w + ox) * 3;
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (ox + i < job.out_w) { o[3 * i] = (uint8_t)val[i][0]; o[3 * i + 1] = (uint8_t)val[i][1]; o[3 * i + 2] = (uint8_t)val[i][2]; }
        } else {
            // planes are R,G,B = source channels 2,1,0
            float* o = out_f32 + tile_base + (int64_t)oy * job.out_w + ox;
            const bool vec = (ox + 3 < job.out_w) && ((((uintptr_t)o) & 15) == 0) && ((plane & 3) == 0);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float4 v;
                v.x = u8_over_255(val[0][2 - c]); v.y = u8_over_255(val[1][2 - c]);
                v.z = u8_over_255(val[2][2 - c]); v.w = u8_over_255(val[3][2 - c]);
                float* oc = o + c * plane;
                if (vec) {
                    *reinterpret_cast<float4*>(oc) = v;
                } else {
                    if (ox + 0 < job.out_w) oc[0] = v.x;
                    if (ox + 1 < job.out_w) oc[1] = v.y;
                    if (ox + 2 < job.out_w) oc[2] = v.z;
                    if (ox + 3 < job.out_w) oc[3] = v.w;
                }
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
    uint32_t* blk2job_dev = nullptr;
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
    std::vector<uint32_t> blk;
    int max_rows = 1, max_span = 16;
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
        j.xtab = j.ytab = 0; j.pad_ = 0;
        if (g.new_w == srcs[t].w && g.new_h == srcs[t].h) j.mode = MODE_COPY;
        else if (srcs[t].w == 2 * g.new_w && srcs[t].h == 2 * g.new_h) j.mode = MODE_AREA2;
        else {
            j.mode = MODE_LINEAR;
            auto kx = std::make_pair(srcs[t].w, g.new_w);
            if (!xkey.count(kx)) {
                xkey[kx] = (int)xt.size();
                std::vector<int> s, a0, a1;
                linear_coeffs(srcs[t].w, g.new_w, true, s, a0, a1);
                for (int d = 0; d < g.new_w; d++) xt.push_back({s[d], (int16_t)a0[d], (int16_t)a1[d]});
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
        // shared-memory need of the worst block of this job
        double sx = (double)srcs[t].w / g.new_w, sy = (double)srcs[t].h / g.new_h;
        int rows = (int)ceil(kTH * std::max(sy, 1e-9)) + 3;
        int span = ((int)ceil(kTW * std::max(sx, 1e-9)) + 3) * 3;
        if (j.mode == MODE_COPY) { rows = kTH; span = kTW * 3; }
        max_rows = std::max(max_rows, std::min(rows, srcs[t].h));
        max_span = std::max(max_span, std::min(span, srcs[t].w * 3));
        const int by = hvb_div_up(g.out_h, kTH), bx = hvb_div_up(g.out_w, kTW);
        if (by > 4095 || bx > 4095) { delete p; hvb_set_error("output too large for the block table"); return HVB_ERR_CAPACITY; }
        for (int y = 0; y < by; y++)
            for (int x = 0; x < bx; x++) blk.push_back(((uint32_t)t << 24) | ((uint32_t)y << 12) | (uint32_t)x);

        p->read_bytes += (int64_t)srcs[t].w * srcs[t].h * 3;
        p->write_bytes += (int64_t)3 * g.out_h * g.out_w * 4;
    }
    // every source byte counted once per frame (tiles overlap): algorithmic read = the frame itself
    p->read_bytes = std::min<int64_t>(p->read_bytes, (int64_t)frame_h * frame_w * 3) * n_frames;
    p->write_bytes *= n_frames;
    p->blocks_per_frame = (int)blk.size();
    p->smem_row_stride = ((max_span + 16 + 15) / 16) * 16 + 16;
    p->smem_bytes = p->smem_row_stride * max_rows;
    if (p->smem_bytes > ctx->max_smem_optin - 1024) {
        delete p;
        hvb_set_error("letterbox block needs %d bytes of shared memory (down-scale factor too large)", p->smem_bytes);
        return HVB_ERR_CAPACITY;
    }
    if ((int64_t)p->blocks_per_frame * n_frames > 0x7fffffffLL) {
        delete p; hvb_set_error("grid too large"); return HVB_ERR_CAPACITY;
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
    if (xt.empty()) xt.push_back({0, 0, 0});
    if (yt.empty()) yt.push_back({0, 0, 0, 0});
    size_t o_jobs = 0;
    size_t o_blk = o_jobs + ((jobs.size() * sizeof(LbJob) + 255) & ~(size_t)255);
    size_t o_x = o_blk + ((blk.size() * sizeof(uint32_t) + 255) & ~(size_t)255);
    size_t o_y = o_x + ((xt.size() * sizeof(XCoef) + 255) & ~(size_t)255);
    size_t total = o_y + yt.size() * sizeof(YCoef);
    std::vector<uint8_t> host(total, 0);
    memcpy(host.data() + o_jobs, jobs.data(), jobs.size() * sizeof(LbJob));
    memcpy(host.data() + o_blk, blk.data(), blk.size() * sizeof(uint32_t));
    memcpy(host.data() + o_x, xt.data(), xt.size() * sizeof(XCoef));
    memcpy(host.data() + o_y, yt.data(), yt.size() * sizeof(YCoef));
    cudaError_t e = cudaMalloc(&p->dev, total);
    if (e != cudaSuccess) { delete p; return hvb_cuda_fail(e, "cudaMalloc(plan)", __FILE__, __LINE__); }
    e = cudaMemcpy(p->dev, host.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p->dev); delete p; return hvb_cuda_fail(e, "cudaMemcpy(plan)", __FILE__, __LINE__); }
    p->jobs_dev = (LbJob*)((uint8_t*)p->dev + o_jobs);
    p->blk2job_dev = (uint32_t*)((uint8_t*)p->dev + o_blk);
    p->xtab_dev = (XCoef*)((uint8_t*)p->dev + o_x);
    p->ytab_dev = (YCoef*)((uint8_t*)p->dev + o_y);
    if (p->smem_bytes > 48 * 1024) {
        HVB_CUDA(cudaFuncSetAttribute(letterbox_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes));
        HVB_CUDA(cudaFuncSetAttribute(letterbox_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes));
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
    const int grid = p->blocks_per_frame * p->n_frames;
    const int64_t frame_bytes = (int64_t)p->frame_h * p->frame_w * 3;
    if (out_u8)
        letterbox_kernel<true><<<grid, kThreads, p->smem_bytes, ctx->stream>>>(
            frames_dev, frame_bytes, p->frame_w * 3, p->jobs_dev, p->blk2job_dev, p->blocks_per_frame, p->xtab_dev,
            p->ytab_dev, p->smem_row_stride, nullptr, out_u8);
    else
        letterbox_kernel<false><<<grid, kThreads, p->smem_bytes, ctx->stream>>>(
            frames_dev, frame_bytes, p->frame_w * 3, p->jobs_dev, p->blk2job_dev, p->blocks_per_frame, p->xtab_dev,
            p->ytab_dev, p->smem_row_stride, out_f32, nullptr);
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

// K2a — YOLOv8 Detect head decode + confidence gate + per-image class-aware NMS + scale_boxes,
// one launch per batch: grid = (anchor chunks, images).  Replaces ultralytics Detect._inference /
// ops.non_max_suppression (torchvision.ops.nms) / ops.scale_boxes reached from
// hockey/main.py:179-186.
//
// Pipeline (256 threads per CTA):
//   1. every CTA scans ONE 1024-anchor chunk of one image: read only the nc class logits per anchor,
//      cheap logit pre-gate (x > logit(thr) - margin) and the exact fp32 sigmoid only for the few logits
//      that pass it, first-max over classes, conf > thr gate; survivors are appended to the image's
//      key list in global scratch  key = score_bits << 32 | (0xFFFFFF - anchor) << 8 | cls.
//      The chunk CTAs of an image count themselves off on a device counter; the LAST one to finish
//      (threadFenceReduction pattern) pulls the list into shared memory and runs steps 2-5, so the
//      latency-bound scan is spread over all SMs while sort + NMS stay on chip.  The append order is
//      arbitrary but keys are unique per image, so the sorted order — and the result — is deterministic.
//   2. bitonic sort of the keys (descending): score order, ties -> lower anchor index first
//      (== torchvision's stable descending sort of the filtered rows)
//   3. decode the DFL box of each survivor only (64 strided reads + 4 softmax-expectations)
//   4. greedy NMS, exactly sequential-greedy but evaluated 32 candidates at a time:
//      (A) all warps test the chunk against the boxes kept so far (kept list split across
//      warps, one __ballot_sync per warp), (B) warp 0 resolves the chunk internally with 32
//      ballot steps; stops as soon as max_det boxes are kept.
//      IoU = inter / (area_i + area_j - inter) in fp32 with IEEE division, suppress iff IoU > thr,
//      boxes offset by cls * 7680 in fp32 BEFORE the IoU (class-aware trick of ultralytics).
//   5. scale_boxes + clip on the kept boxes, written in score order.
// The library is compiled with --fmad=false so none of the mul/add chains are contracted.
#include "hvb_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kCapSmall = 1024;
constexpr int kCapLarge = 8192;
constexpr float kMaxWH = 7680.0f;

struct Levels {
    const float* ptr[3];          // box-bin channels (64 per anchor)
    const float* cls[3];          // class-logit channel 0 (same tensor + 64 channels, or a separate dense tensor)
    int32_t h[3], w[3];
    int64_t bstride[3], cstride[3], astride[3];
    int64_t cls_bstride[3], cls_cstride[3], cls_astride[3];
    int32_t base[4];      // anchor index base per level, base[3] = A
};

// order-preserving map float -> uint32 (handles negative scores)
__device__ __forceinline__ unsigned ordered_bits(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float sigmoidf_exact(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__device__ __forceinline__ void locate(const Levels& L, int a, int& lvl, int& q) {
    lvl = (a >= L.base[1]) + (a >= L.base[2]);
    q = a - L.base[lvl];
}

// Detect._inference for one anchor: DFL softmax-expectation per side, dist2bbox(xywh), x stride.
__device__ __forceinline__ void decode_xywh(const Levels& L, int b, int lvl, int q, float& cx, float& cy, float& w, float& h) {
    const float* p = L.ptr[lvl] + (int64_t)b * L.bstride[lvl] + (int64_t)q * L.astride[lvl];
    const int64_t cs = L.cstride[lvl];
    float dist[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        float v[16];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < 16; k++) { v[k] = __ldg(p + (int64_t)(s * 16 + k) * cs); m = fmaxf(m, v[k]); }
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k++) { v[k] = expf(__fsub_rn(v[k], m)); sum = __fadd_rn(sum, v[k]); }
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k++) acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(v[k], sum), (float)k));
        dist[s] = acc;
    }
    const int gx = q % L.w[lvl], gy = q / L.w[lvl];
    const float ax = (float)gx + 0.5f, ay = (float)gy + 0.5f;
    const float stride = (float)(8 << lvl);
    float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
    cx = __fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.0f), stride);
    cy = __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.0f), stride);
    w = __fmul_rn(__fsub_rn(x2, x1), stride);
    h = __fmul_rn(__fsub_rn(y2, y1), stride);
}

__device__ __forceinline__ bool iou_gt(const float4& a, const float4& b, float thr) {
    // torchvision nms: areas (x2-x1)*(y2-y1); w = max(0, xx2-xx1); ovr = inter/(areaA+areaB-inter)
    float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    float inter = __fmul_rn(w, h);
    float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return ovr > thr;
}

// Descending bitonic sort of n_pow2 64-bit keys in shared memory.
__device__ void bitonic_sort_desc(unsigned long long* keys, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += kThreads) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = keys[i], b = keys[ixj];
                    bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// Greedy NMS over n sorted boxes (already class-offset) in shared memory.  Returns kept count;
// kept positions (indices into the sorted order) are written to s_keptidx.
__device__ int greedy_nms(const float4* s_box, int n, float thr, int max_det, float4* s_keptbox, int* s_keptidx,
                          unsigned* s_sup, int* s_nkept) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) *s_nkept = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int i = c0 + lane;
        const bool valid = i < n;
        const float4 bi = valid ? s_box[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int kept_n = *s_nkept;
        bool sup = false;
        for (int k = warp; k < kept_n; k += kWarps) {
            if (valid && iou_gt(s_keptbox[k], bi, thr)) sup = true;
        }
        unsigned m = __ballot_sync(0xffffffffu, sup);
        if (lane == 0) s_sup[warp] = m;
        __syncthreads();
        if (warp == 0) {
            unsigned supmask = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) supmask |= s_sup[w];
            unsigned alive = __ballot_sync(0xffffffffu, valid) & ~supmask;
            for (int j = 0; j < 31; j++) {
                if (!((alive >> j) & 1u)) continue;           // warp-uniform
                float4 bj;
                bj.x = __shfl_sync(0xffffffffu, bi.x, j); bj.y = __shfl_sync(0xffffffffu, bi.y, j);
                bj.z = __shfl_sync(0xffffffffu, bi.z, j); bj.w = __shfl_sync(0xffffffffu, bi.w, j);
                bool kill = (lane > j) && ((alive >> lane) & 1u) && iou_gt(bj, bi, thr);
                alive &= ~__ballot_sync(0xffffffffu, kill);
            }
            const int pos = kept_n + __popc(alive & ((1u << lane) - 1u));
            if (((alive >> lane) & 1u) && pos < max_det) { s_keptbox[pos] = bi; s_keptidx[pos] = i; }
            if (lane == 0) *s_nkept = min(kept_n + __popc(alive), max_det);
        }
        __syncthreads();
        if (*s_nkept >= max_det) break;
    }
    return *s_nkept;
}

struct NmsSmem {
    unsigned long long* keys;   // [cap]
    float4* box;                // [cap] sorted, class-offset boxes
    float4* keptbox;            // [max_det]
    int* keptidx;               // [max_det]
};

__device__ __forceinline__ NmsSmem carve(uint8_t* smem, int cap, int max_det) {
    NmsSmem s;
    s.keys = reinterpret_cast<unsigned long long*>(smem);
    s.box = reinterpret_cast<float4*>(smem + (size_t)cap * 8);
    s.keptbox = reinterpret_cast<float4*>(smem + (size_t)cap * 24);
    s.keptidx = reinterpret_cast<int*>(smem + (size_t)cap * 24 + (size_t)max_det * 16);
    return s;
}

size_t nms_smem_bytes(int cap, int max_det) { return (size_t)cap * 24 + (size_t)max_det * 20; }

constexpr int kChunk = 1024;                         // anchors scanned per CTA
constexpr int kPerThread = kChunk / kThreads;

__global__ void __launch_bounds__(kThreads)
decode_nms_kernel(Levels L, int nc, float conf_thres, float pre_gate, float iou_thres, int max_det, int agnostic, int cap,
                  const hvb_img_meta* __restrict__ meta, float* __restrict__ out_xyxy, float* __restrict__ out_conf,
                  int32_t* __restrict__ out_cls, int32_t* __restrict__ out_count, const int32_t* __restrict__ only_images,
                  unsigned long long* __restrict__ g_keys /*[images][cap]*/, int32_t* __restrict__ g_ctr /*[images][2]*/) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_count;
    __shared__ int s_nkept;
    __shared__ unsigned s_sup[kWarps];
    NmsSmem S = carve(smem, cap, max_det);

    const int img = blockIdx.y;
    const int b = only_images ? only_images[img] : img;
    unsigned long long* keys_g = g_keys + (size_t)img * cap;
    int32_t* ctr = g_ctr + 2 * img;                   // [0] candidates appended, [1] chunk CTAs finished

    // ---- 1. scan this CTA's chunk of class logits, gate, append to the image's global key list.
    // All loads of the kPerThread anchors are issued before the first append (no sync in between).
    const int A = L.base[3];
    const int a_begin = blockIdx.x * kChunk;
    float best[kPerThread];
    int bestc[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; i++) {
        const int a = a_begin + i * kThreads + threadIdx.x;
        best[i] = -1.f; bestc[i] = 0;
        if (a < A) {
            int lvl, q;
            locate(L, a, lvl, q);
            const float* p = L.cls[lvl] + (int64_t)b * L.cls_bstride[lvl] + (int64_t)q * L.cls_astride[lvl];
            for (int c = 0; c < nc; c++) {
                const float x = __ldg(p + (int64_t)c * L.cls_cstride[lvl]);
                // a logit at or below the pre-gate has sigmoid <= conf_thres, so it can neither pass nor be the max of a
                // passing anchor; the exact sigmoid (expf + IEEE division) is evaluated only above it
                if (x > pre_gate) {
                    const float sg = sigmoidf_exact(x);
                    if (sg > best[i]) { best[i] = sg; bestc[i] = c; }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kPerThread; i++) {
        const int a = a_begin + i * kThreads + threadIdx.x;
        const bool pass = a < A && best[i] > conf_thres;
        const unsigned m = __ballot_sync(0xffffffffu, pass);       // warp-aggregated append
        if (m) {
            const int lane = threadIdx.x & 31;
            int base = 0;
            if (lane == 0) base = atomicAdd(&ctr[0], __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) {
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < cap)
                    keys_g[slot] = ((unsigned long long)__float_as_uint(best[i]) << 32) |
                                   ((unsigned long long)(0xFFFFFFu - (unsigned)a) << 8) | (unsigned)bestc[i];
            }
        }
    }
    // ---- last chunk CTA of the image takes over
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(&ctr[1], 1);
        s_count = (prev == (int)gridDim.x - 1) ? atomicAdd(&ctr[0], 0) : -1;
        if (s_count >= 0) { ctr[0] = 0; ctr[1] = 0; }             // ready for the next launch
    }
    __syncthreads();
    const int total = s_count;
    if (total < 0) return;
    __threadfence();
    if (total > cap) {                       // capacity exceeded: report, host retries with the large tier
        if (threadIdx.x == 0) out_count[meta[b].out_slot] = -1;
        return;
    }
    for (int i = threadIdx.x; i < total; i += kThreads) S.keys[i] = __ldcg(keys_g + i);
    const int n = total;
    int n_pow2 = 1;
    while (n_pow2 < n) n_pow2 <<= 1;
    for (int i = n + threadIdx.x; i < n_pow2; i += kThreads) S.keys[i] = 0ull;
    __syncthreads();

    // ---- 2. sort
    bitonic_sort_desc(S.keys, n_pow2);

    // ---- 3. decode survivors -> xyxy (+ class offset)
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const unsigned long long key = S.keys[i];
        const int a = (int)(0xFFFFFFu - (unsigned)((key >> 8) & 0xFFFFFFu));
        const int c = (int)(key & 0xFFu);
        int lvl, q;
        locate(L, a, lvl, q);
        float cx, cy, w, h;
        decode_xywh(L, b, lvl, q, cx, cy, w, h);
        const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);     // xywh2xyxy
        const float off = agnostic ? 0.f : __fmul_rn((float)c, kMaxWH);
        float4 bx;
        bx.x = __fadd_rn(__fsub_rn(cx, hw), off); bx.y = __fadd_rn(__fsub_rn(cy, hh), off);
        bx.z = __fadd_rn(__fadd_rn(cx, hw), off); bx.w = __fadd_rn(__fadd_rn(cy, hh), off);
        S.box[i] = bx;
    }
    __syncthreads();

    // ---- 4. greedy NMS
    const int nk = greedy_nms(S.box, n, iou_thres, max_det, S.keptbox, S.keptidx, s_sup, &s_nkept);

    // ---- 5. outputs: original (un-offset) box -> scale_boxes -> clip
    const hvb_img_meta mt = meta[b];
    for (int k = threadIdx.x; k < nk; k += kThreads) {
        const int i = S.keptidx[k];
        const unsigned long long key = S.keys[i];
        const int a = (int)(0xFFFFFFu - (unsigned)((key >> 8) & 0xFFFFFFu));
        const int c = (int)(key & 0xFFu);
        const float score = __uint_as_float((unsigned)(key >> 32));
        int lvl, q;
        locate(L, a, lvl, q);
        float cx, cy, w, h;
        decode_xywh(L, b, lvl, q, cx, cy, w, h);
        const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
        float x1 = __fsub_rn(cx, hw), y1 = __fsub_rn(cy, hh), x2 = __fadd_rn(cx, hw), y2 = __fadd_rn(cy, hh);
        x1 = __fdiv_rn(__fsub_rn(x1, mt.pad_x), mt.gain); y1 = __fdiv_rn(__fsub_rn(y1, mt.pad_y), mt.gain);
        x2 = __fdiv_rn(__fsub_rn(x2, mt.pad_x), mt.gain); y2 = __fdiv_rn(__fsub_rn(y2, mt.pad_y), mt.gain);
        x1 = fminf(fmaxf(x1, 0.f), mt.clip_w); x2 = fminf(fmaxf(x2, 0.f), mt.clip_w);
        y1 = fminf(fmaxf(y1, 0.f), mt.clip_h); y2 = fminf(fmaxf(y2, 0.f), mt.clip_h);
        const int64_t o = (int64_t)mt.out_slot * max_det + k;
        out_xyxy[4 * o + 0] = x1; out_xyxy[4 * o + 1] = y1; out_xyxy[4 * o + 2] = x2; out_xyxy[4 * o + 3] = y2;
        out_conf[o] = score;
        out_cls[o] = c;
    }
    if (threadIdx.x == 0) out_count[mt.out_slot] = nk;
}

// Full decode (test hook): thread per anchor, writes float32[batch, 4+nc, A].
__global__ void __launch_bounds__(kThreads)
decode_only_kernel(Levels L, int nc, float* __restrict__ out) {
    const int A = L.base[3];
    const int b = blockIdx.y;
    const int a = blockIdx.x * kThreads + threadIdx.x;
    if (a >= A) return;
    int lvl, q;
    locate(L, a, lvl, q);
    float cx, cy, w, h;
    decode_xywh(L, b, lvl, q, cx, cy, w, h);
    float* o = out + (int64_t)b * (4 + nc) * A + a;
    o[0] = cx; o[(int64_t)A] = cy; o[2 * (int64_t)A] = w; o[3 * (int64_t)A] = h;
    const float* p = L.cls[lvl] + (int64_t)b * L.cls_bstride[lvl] + (int64_t)q * L.cls_astride[lvl];
    for (int c = 0; c < nc; c++) o[(int64_t)(4 + c) * A] = sigmoidf_exact(__ldg(p + (int64_t)c * L.cls_cstride[lvl]));
}

// NMS on caller-provided candidates of one image (test hook for the margin-free keep-set test).
__global__ void __launch_bounds__(kThreads)
nms_only_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ cls, int n,
                float iou_thres, int max_det, int agnostic, int cap, int32_t* __restrict__ out_keep, int32_t* __restrict__ out_count) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_nkept;
    __shared__ unsigned s_sup[kWarps];
    NmsSmem S = carve(smem, cap, max_det);
    int n_pow2 = 1;
    while (n_pow2 < n) n_pow2 <<= 1;
    for (int i = threadIdx.x; i < n_pow2; i += kThreads)
        S.keys[i] = i < n ? (((unsigned long long)ordered_bits(scores[i]) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i)) : 0ull;
    __syncthreads();
    bitonic_sort_desc(S.keys, n_pow2);
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const int src = (int)(0xFFFFFFFFu - (unsigned)(S.keys[i] & 0xFFFFFFFFu));
        const float off = agnostic ? 0.f : __fmul_rn((float)cls[src], kMaxWH);
        float4 bx;
        bx.x = __fadd_rn(boxes[4 * src + 0], off); bx.y = __fadd_rn(boxes[4 * src + 1], off);
        bx.z = __fadd_rn(boxes[4 * src + 2], off); bx.w = __fadd_rn(boxes[4 * src + 3], off);
        S.box[i] = bx;
    }
    __syncthreads();
    const int nk = greedy_nms(S.box, n, iou_thres, max_det, S.keptbox, S.keptidx, s_sup, &s_nkept);
    for (int k = threadIdx.x; k < nk; k += kThreads)
        out_keep[k] = (int)(0xFFFFFFFFu - (unsigned)(S.keys[S.keptidx[k]] & 0xFFFFFFFFu));
    if (threadIdx.x == 0) *out_count = nk;
}

int fill_levels(Levels& L, const float* const level_dev[3], const int32_t level_h[3], const int32_t level_w[3],
                const int64_t batch_stride[3], const int64_t chan_stride[3], const int64_t anchor_stride[3]) {
    int base = 0;
    for (int i = 0; i < 3; i++) {
        if (!level_dev[i] || level_h[i] <= 0 || level_w[i] <= 0) { hvb_set_error("bad level %d", i); return HVB_ERR_ARG; }
        L.ptr[i] = level_dev[i]; L.h[i] = level_h[i]; L.w[i] = level_w[i];
        L.bstride[i] = batch_stride[i]; L.cstride[i] = chan_stride[i]; L.astride[i] = anchor_stride[i];
        // class logits: channels 64.. of the same tensor unless fill_cls() points them at a separate one
        L.cls[i] = level_dev[i] + 64 * chan_stride[i];
        L.cls_bstride[i] = batch_stride[i]; L.cls_cstride[i] = chan_stride[i]; L.cls_astride[i] = anchor_stride[i];
        L.base[i] = base;
        base += level_h[i] * level_w[i];
    }
    L.base[3] = base;
    if (base >= (1 << 24)) { hvb_set_error("more than 2^24 anchors per image"); return HVB_ERR_CAPACITY; }
    return HVB_OK;
}

int fill_cls(Levels& L, const float* const cls_dev[3], const int64_t batch_stride[3], const int64_t chan_stride[3],
             const int64_t anchor_stride[3]) {
    for (int i = 0; i < 3; i++) {
        if (!cls_dev[i]) { hvb_set_error("null class-logit tensor for level %d", i); return HVB_ERR_ARG; }
        L.cls[i] = cls_dev[i];
        L.cls_bstride[i] = batch_stride[i]; L.cls_cstride[i] = chan_stride[i]; L.cls_astride[i] = anchor_stride[i];
    }
    return HVB_OK;
}

// Conservative logit-space version of `sigmoid(x) > conf`: everything at or below the returned value has
// sigmoid(x) <= conf with a margin (1e-3 in logit space) far larger than the fp32 error of the exact sigmoid.
float logit_pre_gate(float conf) {
    if (!(conf > 0.0f)) return -INFINITY;
    if (conf >= 1.0f) return INFINITY;
    const double lg = log((double)conf / (1.0 - (double)conf));
    return (float)(lg - 1e-3 - 1e-5 * fabs(lg));
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) {
        hvb_set_error("NMS needs %zu bytes of shared memory (> 227 KB): lower max_det for images with more than 1024 candidates", bytes);
        return HVB_ERR_CAPACITY;
    }
    if (bytes > 48 * 1024) HVB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return HVB_OK;
}

}  // namespace

extern "C" {

int hvb_nms_capacity(int* out_max_candidates) {
    if (!out_max_candidates) { hvb_set_error("null argument"); return HVB_ERR_ARG; }
    *out_max_candidates = kCapLarge;
    return HVB_OK;
}

static int decode_nms_impl(hvb_ctx* ctx, const float* const level_dev[3], const float* const cls_dev[3],
                           const int32_t level_h[3], const int32_t level_w[3], const int64_t batch_stride[3],
                           const int64_t chan_stride[3], const int64_t anchor_stride[3], const int64_t cls_batch_stride[3],
                           const int64_t cls_chan_stride[3], const int64_t cls_anchor_stride[3], const int32_t* images_dev,
                           int batch, int nc, float conf_thres, float iou_thres, int max_det, int agnostic,
                           const hvb_img_meta* meta_dev, float* out_xyxy_dev, float* out_conf_dev, int32_t* out_cls_dev,
                           int32_t* out_count_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(batch >= 0 && nc >= 1 && nc <= 256, "bad batch/nc");
    HVB_ARG(max_det >= 1 && max_det <= 2048, "max_det out of range [1,2048]");
    if (batch == 0) return HVB_OK;
    HVB_ARG(meta_dev && out_xyxy_dev && out_conf_dev && out_cls_dev && out_count_dev, "null pointer");
    HVB_ARG(batch <= 65535, "more than 65535 images in one launch");
    Levels L;
    HVB_TRY(fill_levels(L, level_dev, level_h, level_w, batch_stride, chan_stride, anchor_stride));
    if (cls_dev) HVB_TRY(fill_cls(L, cls_dev, cls_batch_stride, cls_chan_stride, cls_anchor_stride));
    const int cap = images_dev ? kCapLarge : kCapSmall;           // the retry tier works on a list of images
    const size_t sm = nms_smem_bytes(cap, max_det);
    HVB_TRY(set_smem(decode_nms_kernel, sm));
    unsigned long long* keys = nullptr;
    int32_t* ctr = nullptr;
    HVB_TRY(hvb_k2_work(ctx, batch, cap, &keys, &ctr));
    dim3 grid(hvb_div_up(L.base[3], kChunk), batch);
    decode_nms_kernel<<<grid, kThreads, sm, ctx->stream>>>(L, nc, conf_thres, logit_pre_gate(conf_thres), iou_thres, max_det,
                                                           agnostic, cap, meta_dev, out_xyxy_dev, out_conf_dev, out_cls_dev,
                                                           out_count_dev, images_dev, keys, ctr);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_decode_nms_split(hvb_ctx* ctx, const float* const box_level_dev[3], const float* const cls_level_dev[3],
                         const int32_t level_h[3], const int32_t level_w[3], const int64_t box_batch_stride[3],
                         const int64_t box_chan_stride[3], const int64_t box_anchor_stride[3],
                         const int64_t cls_batch_stride[3], const int64_t cls_chan_stride[3],
                         const int64_t cls_anchor_stride[3], const int32_t* images_dev, int batch, int nc, float conf_thres,
                         float iou_thres, int max_det, int agnostic, const hvb_img_meta* meta_dev, float* out_xyxy_dev,
                         float* out_conf_dev, int32_t* out_cls_dev, int32_t* out_count_dev) {
    if (!cls_level_dev) { hvb_set_error("hvb_decode_nms_split: null class-logit tensors"); return HVB_ERR_ARG; }
    return decode_nms_impl(ctx, box_level_dev, cls_level_dev, level_h, level_w, box_batch_stride, box_chan_stride,
                           box_anchor_stride, cls_batch_stride, cls_chan_stride, cls_anchor_stride, images_dev, batch, nc,
                           conf_thres, iou_thres, max_det, agnostic, meta_dev, out_xyxy_dev, out_conf_dev, out_cls_dev,
                           out_count_dev);
}

int hvb_decode_nms(hvb_ctx* ctx, const float* const level_dev[3], const int32_t level_h[3], const int32_t level_w[3],
                   const int64_t batch_stride[3], const int64_t chan_stride[3], const int64_t anchor_stride[3], int batch,
                   int nc, float conf_thres, float iou_thres, int max_det, int agnostic, const hvb_img_meta* meta_dev,
                   float* out_xyxy_dev, float* out_conf_dev, int32_t* out_cls_dev, int32_t* out_count_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(batch >= 0 && nc >= 1 && nc <= 256, "bad batch/nc");
    HVB_ARG(max_det >= 1 && max_det <= 2048, "max_det out of range [1,2048]");
    if (batch == 0) return HVB_OK;
    HVB_ARG(meta_dev && out_xyxy_dev && out_conf_dev && out_cls_dev && out_count_dev, "null pointer");
    Levels L;
    HVB_TRY(fill_levels(L, level_dev, level_h, level_w, batch_stride, chan_stride, anchor_stride));

    // Tier 1: small shared-memory capacity for every image.
    HVB_ARG(batch <= 65535, "more than 65535 images in one launch");
    size_t sm1 = nms_smem_bytes(kCapSmall, max_det);
    HVB_TRY(set_smem(decode_nms_kernel, sm1));
    unsigned long long* keys = nullptr;
    int32_t* ctr = nullptr;
    HVB_TRY(hvb_k2_work(ctx, batch, kCapSmall, &keys, &ctr));
    dim3 grid(hvb_div_up(L.base[3], kChunk), batch);
    decode_nms_kernel<<<grid, kThreads, sm1, ctx->stream>>>(L, nc, conf_thres, logit_pre_gate(conf_thres), iou_thres, max_det,
                                                            agnostic, kCapSmall, meta_dev, out_xyxy_dev, out_conf_dev,
                                                            out_cls_dev, out_count_dev, nullptr, keys, ctr);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_decode_nms_large(hvb_ctx* ctx, const float* const level_dev[3], const int32_t level_h[3],
                                 const int32_t level_w[3], const int64_t batch_stride[3], const int64_t chan_stride[3],
                                 const int64_t anchor_stride[3], const int32_t* images_dev, int n_images, int nc,
                                 float conf_thres, float iou_thres, int max_det, int agnostic,
                                 const hvb_img_meta* meta_dev, float* out_xyxy_dev, float* out_conf_dev,
                                 int32_t* out_cls_dev, int32_t* out_count_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n_images >= 0 && nc >= 1 && nc <= 256, "bad n_images/nc");
    HVB_ARG(max_det >= 1 && max_det <= 2048, "max_det out of range [1,2048]");
    if (n_images == 0) return HVB_OK;
    HVB_ARG(images_dev && meta_dev && out_xyxy_dev && out_conf_dev && out_cls_dev && out_count_dev, "null pointer");
    Levels L;
    HVB_TRY(fill_levels(L, level_dev, level_h, level_w, batch_stride, chan_stride, anchor_stride));
    HVB_ARG(n_images <= 65535, "more than 65535 images in one launch");
    size_t sm = nms_smem_bytes(kCapLarge, max_det);
    HVB_TRY(set_smem(decode_nms_kernel, sm));
    unsigned long long* keys = nullptr;
    int32_t* ctr = nullptr;
    HVB_TRY(hvb_k2_work(ctx, n_images, kCapLarge, &keys, &ctr));
    dim3 grid(hvb_div_up(L.base[3], kChunk), n_images);
    decode_nms_kernel<<<grid, kThreads, sm, ctx->stream>>>(L, nc, conf_thres, logit_pre_gate(conf_thres), iou_thres, max_det,
                                                           agnostic, kCapLarge, meta_dev, out_xyxy_dev, out_conf_dev,
                                                           out_cls_dev, out_count_dev, images_dev, keys, ctr);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_decode_only(hvb_ctx* ctx, const float* const level_dev[3], const int32_t level_h[3], const int32_t level_w[3],
                    const int64_t batch_stride[3], const int64_t chan_stride[3], const int64_t anchor_stride[3], int batch,
                    int nc, float* out_pred_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(batch >= 0 && nc >= 1 && nc <= 256, "bad batch/nc");
    if (batch == 0) return HVB_OK;
    HVB_ARG(out_pred_dev != nullptr, "null output");
    Levels L;
    HVB_TRY(fill_levels(L, level_dev, level_h, level_w, batch_stride, chan_stride, anchor_stride));
    dim3 grid(hvb_div_up(L.base[3], kThreads), batch);
    decode_only_kernel<<<grid, kThreads, 0, ctx->stream>>>(L, nc, out_pred_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_nms_f32(hvb_ctx* ctx, const float* boxes_dev, const float* scores_dev, const int32_t* cls_dev, int n,
                float iou_thres, int max_det, int agnostic, int32_t* out_keep_idx_dev, int32_t* out_count_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0 && max_det >= 1 && max_det <= 2048, "bad n/max_det");
    HVB_ARG(out_keep_idx_dev && out_count_dev, "null output");
    if (n > kCapLarge) { hvb_set_error("hvb_nms_f32: %d candidates exceed the on-chip capacity %d", n, kCapLarge); return HVB_ERR_CAPACITY; }
    if (n == 0) { HVB_CUDA(cudaMemsetAsync(out_count_dev, 0, sizeof(int32_t), ctx->stream)); return HVB_OK; }
    HVB_ARG(boxes_dev && scores_dev && (agnostic || cls_dev), "null input");
    int cap = n <= kCapSmall ? kCapSmall : kCapLarge;
    size_t sm = nms_smem_bytes(cap, max_det);
    HVB_TRY(set_smem(nms_only_kernel, sm));
    nms_only_kernel<<<1, kThreads, sm, ctx->stream>>>(boxes_dev, scores_dev, cls_dev, n, iou_thres, max_det, agnostic, cap,
                                                      out_keep_idx_dev, out_count_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

}  // extern "C"

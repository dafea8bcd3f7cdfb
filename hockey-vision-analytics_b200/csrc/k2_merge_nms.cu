// K2b — cross-slice merge: sv.move_detections + Detections.merge + Detections.with_nms
// (box_non_max_suppression / box_iou_batch, float64, keep mask in INPUT order, class-aware unless
// class_agnostic) — SURVEY.md App. B2; reference documentation README.md:25, CLAUDE.md:55.
//
// hvb_gather_tiles: per-tile K2a results (fixed max_det rows per slot) -> per-frame merged lists in
//   slicer tile order, boxes promoted to float64 and moved by the tile offset (numpy promotes
//   float32 boxes + int64 offsets to float64), plus the segment table.
// hvb_merge_nms: one CTA per segment (frame): sort by score (descending; ties -> higher input
//   index first, i.e. np.flip of a stable ascending argsort), then the same chunked greedy scheme as
//   K2a in float64 with the category-equality condition, writing keep flags at the input positions.
#include "hvb_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kCapSmall = 512;
constexpr int kCapLarge = 4096;

struct DBox { double x1, y1, x2, y2; };

__device__ __forceinline__ unsigned ordered_bits(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ bool iou_gt_f64(const DBox& a, const DBox& b, double thr) {
    double area_a = __dmul_rn(__dsub_rn(a.x2, a.x1), __dsub_rn(a.y2, a.y1));
    double area_b = __dmul_rn(__dsub_rn(b.x2, b.x1), __dsub_rn(b.y2, b.y1));
    double w = fmax(__dsub_rn(fmin(a.x2, b.x2), fmax(a.x1, b.x1)), 0.0);
    double h = fmax(__dsub_rn(fmin(a.y2, b.y2), fmax(a.y1, b.y1)), 0.0);
    double inter = __dmul_rn(w, h);
    double iou = __ddiv_rn(inter, __dsub_rn(__dadd_rn(area_a, area_b), inter));
    return iou > thr;     // NaN (0/0) compares false == nan_to_num -> 0
}

__global__ void __launch_bounds__(kThreads)
merge_nms_kernel(const double* __restrict__ xyxy, const float* __restrict__ conf, const int32_t* __restrict__ cls,
                 const int32_t* __restrict__ seg_offsets, double thr, int agnostic, int cap, int min_n,
                 uint8_t* __restrict__ keep) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_nkept;
    __shared__ unsigned s_sup[kWarps];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);                // [cap]
    DBox* box = reinterpret_cast<DBox*>(smem + (size_t)cap * 8);                           // [cap] sorted
    int* cat = reinterpret_cast<int*>(smem + (size_t)cap * 40);                            // [cap] sorted
    int* kidx = reinterpret_cast<int*>(smem + (size_t)cap * 44);                           // [cap] kept -> sorted index

    const int lo = seg_offsets[blockIdx.x], n = seg_offsets[blockIdx.x + 1] - lo;
    if (n <= min_n) return;          // tiered launches: this launch handles segments with min_n < n <= cap
    if (n > cap) {
        // left to the next tier; beyond the last tier the whole segment is marked 0xFF and the host layer raises
        if (cap == kCapLarge)
            for (int i = threadIdx.x; i < n; i += kThreads) keep[lo + i] = 0xFF;
        return;
    }
    int n_pow2 = 1;
    while (n_pow2 < n) n_pow2 <<= 1;
    for (int i = threadIdx.x; i < n_pow2; i += kThreads)
        keys[i] = i < n ? (((unsigned long long)ordered_bits(conf[lo + i]) << 32) | (unsigned)i) : 0ull;
    __syncthreads();
    for (int k = 2; k <= n_pow2; k <<= 1)
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
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const int src = lo + (int)(keys[i] & 0xFFFFFFFFu);
        box[i] = DBox{xyxy[4 * (int64_t)src], xyxy[4 * (int64_t)src + 1], xyxy[4 * (int64_t)src + 2], xyxy[4 * (int64_t)src + 3]};
        cat[i] = (agnostic || !cls) ? 0 : cls[src];
        keep[src] = 0;
    }
    if (threadIdx.x == 0) s_nkept = 0;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int i = c0 + lane;
        const bool valid = i < n;
        DBox bi = valid ? box[i] : DBox{0, 0, 0, 0};
        const int ci = valid ? cat[i] : -1;
        const int kept_n = s_nkept;
        bool sup = false;
        for (int k = warp; k < kept_n; k += kWarps)
            if (valid && cat[kidx[k]] == ci && iou_gt_f64(box[kidx[k]], bi, thr)) sup = true;
        unsigned m = __ballot_sync(0xffffffffu, sup);
        if (lane == 0) s_sup[warp] = m;
        __syncthreads();
        if (warp == 0) {
            unsigned supmask = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) supmask |= s_sup[w];
            unsigned alive = __ballot_sync(0xffffffffu, valid) & ~supmask;
            for (int j = 0; j < 31; j++) {
                if (!((alive >> j) & 1u)) continue;
                DBox bj;
                bj.x1 = __shfl_sync(0xffffffffu, bi.x1, j); bj.y1 = __shfl_sync(0xffffffffu, bi.y1, j);
                bj.x2 = __shfl_sync(0xffffffffu, bi.x2, j); bj.y2 = __shfl_sync(0xffffffffu, bi.y2, j);
                const int cj = __shfl_sync(0xffffffffu, ci, j);
                bool kill = (lane > j) && ((alive >> lane) & 1u) && cj == ci && iou_gt_f64(bj, bi, thr);
                alive &= ~__ballot_sync(0xffffffffu, kill);
            }
            if ((alive >> lane) & 1u) {
                const int pos = kept_n + __popc(alive & ((1u << lane) - 1u));
                kidx[pos] = i;
                keep[lo + (int)(keys[i] & 0xFFFFFFFFu)] = 1;
            }
            if (lane == 0) s_nkept = kept_n + __popc(alive);
        }
        __syncthreads();
    }
}

// One CTA: exclusive scan of per-slot counts (frame-major slots), then scatter rows.
__global__ void __launch_bounds__(1024)
gather_tiles_kernel(const float* __restrict__ xyxy, const float* __restrict__ conf, const int32_t* __restrict__ cls,
                    const int32_t* __restrict__ count, const float* __restrict__ slot_off_xy, int n_slots,
                    int slots_per_frame, int max_det, double* __restrict__ out_xyxy, float* __restrict__ out_conf,
                    int32_t* __restrict__ out_cls, int32_t* __restrict__ out_slot, int32_t* __restrict__ seg_offsets,
                    int32_t* __restrict__ slot_base) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int s0 = 0; s0 < n_slots; s0 += 1024) {
        const int s = s0 + threadIdx.x;
        const int c = s < n_slots ? max(count[s], 0) : 0;
        int v = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += t; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int excl = s_carry + (warp ? s_warp[warp - 1] : 0) + v - c;
        if (s < n_slots) {
            slot_base[s] = excl;
            if (s % slots_per_frame == 0) seg_offsets[s / slots_per_frame] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) seg_offsets[n_slots / slots_per_frame] = s_carry;
    __syncthreads();
    // scatter: one warp per slot
    for (int s = warp; s < n_slots; s += 32) {
        const int c = max(count[s], 0), base = slot_base[s];
        const double ox = (double)slot_off_xy[2 * s], oy = (double)slot_off_xy[2 * s + 1];
        for (int k = lane; k < c; k += 32) {
            const int64_t src = (int64_t)s * max_det + k;
            const int64_t dst = base + k;
            out_xyxy[4 * dst + 0] = (double)xyxy[4 * src + 0] + ox;
            out_xyxy[4 * dst + 1] = (double)xyxy[4 * src + 1] + oy;
            out_xyxy[4 * dst + 2] = (double)xyxy[4 * src + 2] + ox;
            out_xyxy[4 * dst + 3] = (double)xyxy[4 * src + 3] + oy;
            out_conf[dst] = conf[src];
            out_cls[dst] = cls[src];
            if (out_slot) out_slot[dst] = s;
        }
    }
}

size_t merge_smem(int cap) { return (size_t)cap * 48; }

}  // namespace

extern "C" {

int hvb_merge_nms(hvb_ctx* ctx, const double* xyxy_dev, const float* conf_dev, const int32_t* cls_dev,
                  const int32_t* seg_offsets_dev, int n_segments, int n_total, double iou_thres, int class_agnostic,
                  uint8_t* out_keep_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n_segments >= 0 && n_total >= 0, "negative sizes");
    if (n_segments == 0 || n_total == 0) return HVB_OK;
    HVB_ARG(xyxy_dev && conf_dev && seg_offsets_dev && out_keep_dev, "null pointer");
    // The tier (shared memory per CTA, bitonic pad) follows the SEGMENT length, which only the device knows when the
    // caller passes capacities (n_total = slots * max_det on the no-sync path): every segment first meets the 512-row
    // tier (24 KB, several CTAs per SM); only if n_total allows longer segments a second launch with the 4096-row tier
    // (192 KB) picks up the ones the first skipped, and its other CTAs exit at once.
    merge_nms_kernel<<<n_segments, kThreads, merge_smem(kCapSmall), ctx->stream>>>(
        xyxy_dev, conf_dev, cls_dev, seg_offsets_dev, iou_thres, class_agnostic, kCapSmall, 0, out_keep_dev);
    HVB_LAUNCHED(ctx);
    if (n_total > kCapSmall) {
        const size_t sm = merge_smem(kCapLarge);
        HVB_CUDA(cudaFuncSetAttribute(merge_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        merge_nms_kernel<<<n_segments, kThreads, sm, ctx->stream>>>(
            xyxy_dev, conf_dev, cls_dev, seg_offsets_dev, iou_thres, class_agnostic, kCapLarge, kCapSmall, out_keep_dev);
        HVB_LAUNCHED(ctx);
    }
    return HVB_OK;
}

int hvb_gather_tiles(hvb_ctx* ctx, const float* xyxy_dev, const float* conf_dev, const int32_t* cls_dev,
                     const int32_t* count_dev, const float* slot_off_xy_dev, int n_slots, int slots_per_frame, int max_det,
                     double* out_xyxy_dev, float* out_conf_dev, int32_t* out_cls_dev, int32_t* out_slot_dev,
                     int32_t* out_seg_offsets_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n_slots >= 0 && slots_per_frame > 0 && max_det > 0, "bad sizes");
    HVB_ARG(n_slots % slots_per_frame == 0, "n_slots must be a multiple of slots_per_frame");
    HVB_ARG(out_seg_offsets_dev != nullptr, "null segment table");
    if (n_slots == 0) { HVB_CUDA(cudaMemsetAsync(out_seg_offsets_dev, 0, sizeof(int32_t), ctx->stream)); return HVB_OK; }
    HVB_ARG(xyxy_dev && conf_dev && cls_dev && count_dev && slot_off_xy_dev && out_xyxy_dev && out_conf_dev && out_cls_dev,
            "null pointer");
    int32_t* slot_base = nullptr;
    HVB_TRY(hvb_scratch2(ctx, (size_t)n_slots * sizeof(int32_t), (void**)&slot_base));
    gather_tiles_kernel<<<1, 1024, 0, ctx->stream>>>(xyxy_dev, conf_dev, cls_dev, count_dev, slot_off_xy_dev, n_slots,
                                                     slots_per_frame, max_det, out_xyxy_dev, out_conf_dev, out_cls_dev,
                                                     out_slot_dev, out_seg_offsets_dev, slot_base);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_merge_nms_host(hvb_ctx* ctx, const double* xyxy_host, const float* conf_host, const int32_t* cls_host, int n,
                       double iou_thres, int class_agnostic, uint8_t* out_keep_host) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n >= 0, "n < 0");
    if (n == 0) return HVB_OK;
    HVB_ARG(xyxy_host && conf_host && out_keep_host, "null pointer");
    if (n > kCapLarge) { hvb_set_error("hvb_merge_nms_host: %d detections exceed the on-chip capacity %d", n, kCapLarge); return HVB_ERR_CAPACITY; }
    const size_t o_conf = (size_t)n * 32, o_cls = o_conf + (size_t)n * 4, o_seg = o_cls + (size_t)n * 4, o_keep = o_seg + 16;
    uint8_t* d = nullptr;
    HVB_TRY(hvb_scratch(ctx, o_keep + n, (void**)&d));
    const int32_t seg[2] = {0, n};
    HVB_CUDA(cudaMemcpyAsync(d, xyxy_host, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
    HVB_CUDA(cudaMemcpyAsync(d + o_conf, conf_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (cls_host) HVB_CUDA(cudaMemcpyAsync(d + o_cls, cls_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    HVB_CUDA(cudaMemcpyAsync(d + o_seg, seg, sizeof(seg), cudaMemcpyHostToDevice, ctx->stream));
    HVB_TRY(hvb_merge_nms(ctx, (const double*)d, (const float*)(d + o_conf), cls_host ? (const int32_t*)(d + o_cls) : nullptr,
                          (const int32_t*)(d + o_seg), 1, n, iou_thres, class_agnostic || !cls_host, d + o_keep));
    HVB_CUDA(cudaMemcpyAsync(out_keep_host, d + o_keep, n, cudaMemcpyDeviceToHost, ctx->stream));
    HVB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HVB_OK;
}

}  // extern "C"

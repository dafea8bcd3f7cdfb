// K7 — device ByteTrack: sv.ByteTrack.update_with_detections (hockey/main.py:162-168, 207-211, 228, 265) for a chunk of
// frames of every clip in ONE launch, fed directly with K2a's per-image detections (no host round trip between the
// detector and the tracker).  SURVEY.md §8(f) rank 1.
//
// Launch shape: one CTA of one warp per clip; the warp walks the clip's frames in order (tracking is sequential per
// clip) and clips run concurrently on different SMs.  Per-frame scratch (detection lists, assignment
// solver arrays, staged boxes: sizeof(BtWork) ~ 64 KB) lives in shared memory; the persistent tracker state (BtClip)
// and the cost matrices live in global memory (L1/L2-resident: a 12 x 12 problem is 1.2 KB).  The algorithm is in
// k7_bytetrack_core.h, shared with a g++ test build.
//
// Latency-bound by construction (dependent fp64 chains of one warp); the figure of merit is clip-frames per second,
// not bytes: see DESIGN.md §4.
#include "hvb_common.cuh"
#include "k7_bytetrack_core.h"

struct hvb_bytetrack {
    int n_clips = 0;
    BtParams params;
    BtClip* clips_dev = nullptr;
    double* cost_dev = nullptr;         // [n_clips][2][BT_N * BT_N]
};

namespace {

__global__ void __launch_bounds__(32)
bytetrack_reset_kernel(BtClip* clips) { bt_reset(&clips[blockIdx.x]); }

__global__ void __launch_bounds__(32)
bytetrack_kernel(BtClip* __restrict__ clips, BtParams p, const float* __restrict__ xyxy, const float* __restrict__ conf,
                 const int32_t* __restrict__ cls, const int32_t* __restrict__ count, int n_frames, int max_det,
                 int64_t clip_stride, int64_t frame_stride, int seq, double* __restrict__ cost_all, int32_t* __restrict__ out_row,
                 int32_t* __restrict__ out_tid, int32_t* __restrict__ out_count) {
    extern __shared__ __align__(16) uint8_t smem[];
    BtWork* w = reinterpret_cast<BtWork*>(smem);
    BtClip* c = &clips[blockIdx.x];
    double* cost = cost_all + (size_t)blockIdx.x * 2 * BT_N * BT_N;
    double* costT = cost + (size_t)BT_N * BT_N;
    const int lane = threadIdx.x & 31;
    // transactional per chunk: if the detector flagged one of this clip's frames (count < 0: candidate overflow, the
    // host is going to redo those frames with the large tier) the tracker state is left untouched and every frame of
    // the chunk reports -2, so the chunk can be stepped again after the retry.  Chunks carry a sequence number and a
    // clip only accepts the one it expects next: chunks that were queued behind a rejected one are rejected too (-2)
    // and resubmitted by the host in order.
    int bad = 0;
    for (int f = lane; f < n_frames; f += 32) bad |= count[blockIdx.x * clip_stride + f * frame_stride] < 0;
    bad = __any_sync(0xffffffffu, bad) || seq != c->next_seq;
    if (bad || c->overflow) {
        for (int f = lane; f < n_frames; f += 32) out_count[blockIdx.x * clip_stride + f * frame_stride] = bad ? -2 : -1;
        return;
    }
    __syncwarp();
    if (lane == 0) c->next_seq = seq + 1;
    __syncwarp();
    for (int f = 0; f < n_frames; f++) {
        const int64_t img = blockIdx.x * clip_stride + f * frame_stride;
        const int n_rows = min(count[img], max_det);
        const int kept = bt_update(c, w, p, xyxy + img * max_det * 4, conf + img * max_det, cls ? cls + img * max_det : nullptr,
                                   n_rows, cost, costT, out_row + img * max_det, out_tid + img * max_det);
        if (lane == 0) out_count[img] = kept;
        __syncwarp();
        if (kept < 0) {                 // capacity exceeded: the remaining frames of the chunk report -1 as well
            for (int g = f + 1 + lane; g < n_frames; g += 32) out_count[blockIdx.x * clip_stride + g * frame_stride] = -1;
            return;
        }
    }
}

}  // namespace

extern "C" {

int hvb_bytetrack_create(hvb_ctx* ctx, int n_clips, double track_activation_threshold, double det_threshold,
                         double minimum_matching_threshold, int max_time_lost, int minimum_consecutive_frames,
                         hvb_bytetrack** out_tracker) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(out_tracker != nullptr, "null out_tracker");
    *out_tracker = nullptr;
    HVB_ARG(n_clips >= 1 && n_clips <= 65535, "n_clips out of range");
    HVB_ARG(minimum_consecutive_frames >= 1 && max_time_lost >= 0, "bad tracker parameters");
    hvb_bytetrack* t = new hvb_bytetrack();
    t->n_clips = n_clips;
    t->params.match_thr = minimum_matching_threshold;
    t->params.det_thr = det_threshold;
    t->params.act_thr = (float)track_activation_threshold;
    t->params.max_time_lost = max_time_lost;
    t->params.min_consec = minimum_consecutive_frames;
    t->params.min_conf = -INFINITY;
    t->params.class_mask = 0xFFFFFFFFu;
    cudaError_t e = cudaMalloc(&t->clips_dev, (size_t)n_clips * sizeof(BtClip));
    if (e == cudaSuccess) e = cudaMalloc(&t->cost_dev, (size_t)n_clips * 2 * BT_N * BT_N * sizeof(double));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bytetrack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BtWork));
    if (e != cudaSuccess) {
        if (t->clips_dev) cudaFree(t->clips_dev);
        if (t->cost_dev) cudaFree(t->cost_dev);
        delete t;
        return hvb_cuda_fail(e, "hvb_bytetrack_create", __FILE__, __LINE__);
    }
    bytetrack_reset_kernel<<<n_clips, 32, 0, ctx->stream>>>(t->clips_dev);
    HVB_LAUNCHED(ctx);
    *out_tracker = t;
    return HVB_OK;
}

int hvb_bytetrack_destroy(hvb_ctx* ctx, hvb_bytetrack* tracker) {
    if (!tracker) return HVB_OK;
    HVB_CHECK_CTX(ctx);
    HVB_CUDA(cudaDeviceSynchronize());
    if (tracker->clips_dev) cudaFree(tracker->clips_dev);
    if (tracker->cost_dev) cudaFree(tracker->cost_dev);
    delete tracker;
    return HVB_OK;
}

int hvb_bytetrack_reset(hvb_ctx* ctx, hvb_bytetrack* tracker) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(tracker != nullptr, "null tracker");
    bytetrack_reset_kernel<<<tracker->n_clips, 32, 0, ctx->stream>>>(tracker->clips_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_bytetrack_update(hvb_ctx* ctx, hvb_bytetrack* tracker, const float* xyxy_dev, const float* conf_dev,
                         const int32_t* cls_dev, const int32_t* count_dev, int n_frames, int max_det, int64_t clip_stride,
                         int64_t frame_stride, float min_conf, uint32_t class_mask, int seq, int32_t* out_row_dev,
                         int32_t* out_tid_dev, int32_t* out_count_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(tracker != nullptr, "null tracker");
    HVB_ARG(n_frames >= 0 && max_det >= 1, "bad sizes");
    if (max_det > BT_D) { hvb_set_error("hvb_bytetrack_update: max_det %d exceeds the tracker's %d detections per frame", max_det, BT_D); return HVB_ERR_CAPACITY; }
    if (n_frames == 0) return HVB_OK;
    HVB_ARG(xyxy_dev && conf_dev && count_dev && out_row_dev && out_tid_dev && out_count_dev, "null pointer");
    HVB_ARG(cls_dev != nullptr || class_mask == 0xFFFFFFFFu, "class_mask needs class ids");
    BtParams p = tracker->params;
    p.min_conf = min_conf;
    p.class_mask = class_mask;
    bytetrack_kernel<<<tracker->n_clips, 32, sizeof(BtWork), ctx->stream>>>(
        tracker->clips_dev, p, xyxy_dev, conf_dev, cls_dev, count_dev, n_frames, max_det, clip_stride, frame_stride, seq,
        tracker->cost_dev, out_row_dev, out_tid_dev, out_count_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

}  // extern "C"

// K4b — ByteTrack association cost: matching.iou_distance (1 - box_iou_batch) and
// matching.fuse_score of supervision's ByteTrack, reached every frame from
// hockey/main.py:228,265 (sv.ByteTrack.update_with_detections).  float64 like numpy; with
// --fmad=false the arithmetic is operation-for-operation the numpy expression, so results are
// bit-identical to box_iou_batch.  Batched over independent problems (one per clip / association).
#include "hvb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
iou_cost_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ scores,
                const int32_t* __restrict__ a_off, const int32_t* __restrict__ b_off, const int64_t* __restrict__ out_off,
                int flags, double* __restrict__ out) {
    const int p = blockIdx.y;
    const int a0 = a_off[p], na = a_off[p + 1] - a0;
    const int b0 = b_off[p], nb = b_off[p + 1] - b0;
    const int64_t total = (int64_t)na * nb;
    double* o = out + out_off[p];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / nb), d = (int)(i % nb);
        const double* ba = a + 4 * (int64_t)(a0 + t);
        const double* bb = b + 4 * (int64_t)(b0 + d);
        const double ax1 = ba[0], ay1 = ba[1], ax2 = ba[2], ay2 = ba[3];
        const double bx1 = bb[0], by1 = bb[1], bx2 = bb[2], by2 = bb[3];
        // numpy computes the area of a float32 box array in float32 before promoting (flags bit 0 / 1)
        const double area_a = (flags & 1) ? (double)__fmul_rn(__fsub_rn((float)ax2, (float)ax1), __fsub_rn((float)ay2, (float)ay1))
                                          : __dmul_rn(__dsub_rn(ax2, ax1), __dsub_rn(ay2, ay1));
        const double area_b = (flags & 2) ? (double)__fmul_rn(__fsub_rn((float)bx2, (float)bx1), __fsub_rn((float)by2, (float)by1))
                                          : __dmul_rn(__dsub_rn(bx2, bx1), __dsub_rn(by2, by1));
        const double w = fmax(__dsub_rn(fmin(ax2, bx2), fmax(ax1, bx1)), 0.0);
        const double h = fmax(__dsub_rn(fmin(ay2, by2), fmax(ay1, by1)), 0.0);
        const double inter = __dmul_rn(w, h);
        double iou = __ddiv_rn(inter, __dsub_rn(__dadd_rn(area_a, area_b), inter));
        if (isnan(iou)) iou = 0.0;                                        // np.nan_to_num
        else if (isinf(iou)) iou = iou > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
        double cost = __dsub_rn(1.0, iou);                                // iou_distance
        if (scores) cost = __dsub_rn(1.0, __dmul_rn(__dsub_rn(1.0, cost), scores[b0 + d]));   // fuse_score
        o[i] = cost;
    }
}

}  // namespace

extern "C" {

int hvb_iou_cost(hvb_ctx* ctx, const double* a_dev, const double* b_dev, const double* scores_dev,
                 const int32_t* a_off_dev, const int32_t* b_off_dev, const int64_t* out_off_dev, int n_problems,
                 int max_na, int max_nb, int flags, double* out_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(n_problems >= 0 && max_na >= 0 && max_nb >= 0, "negative sizes");
    if (n_problems == 0 || max_na == 0 || max_nb == 0) return HVB_OK;
    HVB_ARG(n_problems <= 65535, "too many problems in one call");
    if ((flags & 3) == 3) { hvb_set_error("hvb_iou_cost: both sides float32 is not a ByteTrack case (numpy would compute entirely in float32)"); return HVB_ERR_UNSUPPORTED; }
    HVB_ARG(a_dev && b_dev && a_off_dev && b_off_dev && out_off_dev && out_dev, "null pointer");
    const int64_t cells = (int64_t)max_na * max_nb;
    int gx = (int)((cells + 255) / 256);
    if (gx > ctx->sm_count * 8) gx = ctx->sm_count * 8;
    dim3 grid(gx, n_problems);
    iou_cost_kernel<<<grid, 256, 0, ctx->stream>>>(a_dev, b_dev, scores_dev, a_off_dev, b_off_dev, out_off_dev, flags, out_dev);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int hvb_iou_cost_host(hvb_ctx* ctx, const double* a_host, int na, const double* b_host, int nb, const double* scores_host,
                      int flags, double* out_host) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(na >= 0 && nb >= 0, "negative sizes");
    if (na == 0 || nb == 0) return HVB_OK;
    HVB_ARG(a_host && b_host && out_host, "null pointer");
    const size_t o_b = (size_t)na * 32, o_s = o_b + (size_t)nb * 32, o_meta = o_s + (size_t)nb * 8;
    const size_t o_out = (o_meta + 64 + 255) & ~(size_t)255;
    uint8_t* d = nullptr;
    HVB_TRY(hvb_scratch(ctx, o_out + (size_t)na * nb * 8, (void**)&d));
    struct { int32_t a_off[2]; int32_t b_off[2]; int64_t out_off[1]; } meta = {{0, na}, {0, nb}, {0}};
    HVB_CUDA(cudaMemcpyAsync(d, a_host, (size_t)na * 32, cudaMemcpyHostToDevice, ctx->stream));
    HVB_CUDA(cudaMemcpyAsync(d + o_b, b_host, (size_t)nb * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (scores_host) HVB_CUDA(cudaMemcpyAsync(d + o_s, scores_host, (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
    HVB_CUDA(cudaMemcpyAsync(d + o_meta, &meta, sizeof(meta), cudaMemcpyHostToDevice, ctx->stream));
    HVB_TRY(hvb_iou_cost(ctx, (const double*)d, (const double*)(d + o_b), scores_host ? (const double*)(d + o_s) : nullptr,
                         (const int32_t*)(d + o_meta), (const int32_t*)(d + o_meta + 8), (const int64_t*)(d + o_meta + 16), 1,
                         na, nb, flags, (double*)(d + o_out)));
    HVB_CUDA(cudaMemcpyAsync(out_host, d + o_out, (size_t)na * nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HVB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HVB_OK;
}

}  // extern "C"

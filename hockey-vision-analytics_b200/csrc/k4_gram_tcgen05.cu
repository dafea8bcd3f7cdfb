// K4a (tensor-core part) — G = X . X^T on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// This is the Gram contraction inside the RBF affinity of SpectralClustering(affinity='rbf')
// (sklearn euclidean_distances: -2 X.X^T + |x|^2 + |y|^2), reached from
// hockey/common/team_hybrid.py:185-193.  X is float64 [N, D] (standardised features, D = 625/627).
//
// Precision: tcgen05 has no fp64 kind, so each element is split x = hi + lo with hi, lo both
// TF32-representable (11 significant bits each) and the product is formed as
//   G = hi.hi^T + hi.lo^T + lo.hi^T        (lo.lo^T ~ 2^-22 relative is dropped)
// i.e. ONE kind::tf32 GEMM with K' = 3*Dp, accumulated in fp32 in TMEM.  k4_affinity.cu then
// recomputes in float64 the few pairs whose affinity does not underflow.
//
// Kernel shape (one CTA per 128x128 output tile ON OR ABOVE THE DIAGONAL — G is symmetric, the mirrored tile is
// written by the same epilogue, which halves the tensor work: N=2000 is 136 tiles = one wave of the 148 SMs;
// cta_group::1):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d loads of a 128x32-float A tile and B tile per
//            stage (128-byte swizzle), completion on an mbarrier (expect_tx)
//   warp 1   MMA issuer: one thread issues 4 tcgen05.mma.kind::tf32 (M128,N128,K8) per stage,
//            tcgen05.commit releases the stage / signals the epilogue
//   warp 2   TMEM allocator (128 columns)
//   warps 4-7 epilogue: tcgen05.ld 32x32b.x32 (each warp owns its 32-lane TMEM quadrant) -> global
// Every mbarrier wait is bounded: a protocol bug traps instead of hanging the GPU.
#include "hvb_common.cuh"

#include <cuda.h>

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 32;        // tile; kBK floats = 128 bytes = one swizzle atom
// 3 stages = 96 KB of operand tiles: two CTAs co-reside per SM, so one CTA's epilogue overlaps the other's main loop.
// Measured at N=8192: 297 us with 3 stages, 389 us with 4 or 6 (one CTA per SM).
#ifndef HVB_GRAM_STAGES
#define HVB_GRAM_STAGES 3
#endif
constexpr int kStages = HVB_GRAM_STAGES;
constexpr int kTileBytes = kBM * kBK * 4;             // 16 KB per operand tile
constexpr int kThreads = 256;
constexpr uint32_t kTmemCols = 128;
constexpr unsigned kSpinLimit = 200u * 1000u * 1000u;

struct SharedCtl {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full;
    uint32_t tmem_base;
    uint32_t pad_;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (unsigned spin = 0;; spin++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (spin > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)1 << 16;                               // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}

constexpr uint32_t kIdesc = (1u << 4)                     // D format F32
                            | (2u << 7) | (2u << 10)      // A, B format TF32; both K-major (bits 15,16 = 0)
                            | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

__global__ void __launch_bounds__(kThreads, 1)
gram_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ G, int n, int dp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B needs 1024 B
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kTileBytes;
    SharedCtl* ctl = (SharedCtl*)(smem + 2 * kStages * kTileBytes);

    // G is symmetric: only tiles on or above the diagonal are computed, each one is also stored transposed
    // (the whole CTA leaves before touching any barrier or TMEM, so this is safe)
    if (blockIdx.x < blockIdx.y) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bm = blockIdx.y * kBM, bn = blockIdx.x * kBN;
    const bool mirror = bm != bn;
    const int kb_per_seg = dp / kBK, num_kb = 3 * kb_per_seg;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
        mbar_init(&ctl->tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = ctl->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            // A' = [hi | hi | lo], B' = [hi | lo | hi] addressed inside the single [N, 2*Dp] (hi | lo) array
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                mbar_wait(&ctl->empty[s], phase ^ 1);
                const int seg = kb / kb_per_seg, kk = (kb % kb_per_seg) * kBK;
                const int col_a = (seg == 2 ? dp : 0) + kk;
                const int col_b = (seg == 1 ? dp : 0) + kk;
                mbar_expect_tx(&ctl->full[s], 2 * kTileBytes);
                tma_load_2d(smem_a + s * kTileBytes, &tmap, &ctl->full[s], col_a, bm);
                tma_load_2d(smem_b + s * kTileBytes, &tmap, &ctl->full[s], col_b, bn);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                mbar_wait(&ctl->full[s], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = make_smem_desc(smem_u32(smem_a + s * kTileBytes));
                const uint64_t db = make_smem_desc(smem_u32(smem_b + s * kTileBytes));
#pragma unroll
                for (int k = 0; k < kBK / 8; k++)       // UMMA_K = 8 for tf32: advance 32 bytes (>>4 = 2) inside the atom
                    umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, (kb | k) ? 1u : 0u);
                umma_commit(&ctl->empty[s]);            // frees the smem stage when these MMAs retire
            }
            umma_commit(&ctl->tmem_full);               // accumulator complete
        }
    } else if (warp >= 4) {
        const int q = warp - 4;                         // TMEM lane quadrant == warp_id % 4
        mbar_wait(&ctl->tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = bm + q * 32 + lane;
        const bool vec_ok = (n & 3) == 0;
#pragma unroll 1
        for (int c = 0; c < kBN; c += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < n) {
                float* o = G + (int64_t)row * n + bn + c;
                if (vec_ok && bn + c + 32 <= n) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                        __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (bn + c + j < n) o[j] = __uint_as_float(v[j]);
                }
                if (mirror) {                            // G[col][row]: lanes hold consecutive rows -> coalesced 128-byte stores
                    float* t = G + (int64_t)(bn + c) * n + row;
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (bn + c + j < n) t[(int64_t)j * n] = __uint_as_float(v[j]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// x (float64 [N,D]) -> H float32 [N, 2*Dp]:  H[:, :Dp] = tf32(x),  H[:, Dp:] = tf32(x - hi); zero padded.
__global__ void __launch_bounds__(256)
split_tf32_kernel(const double* __restrict__ x, int n, int d, int dp, float* __restrict__ h) {
    const int64_t total = (int64_t)n * dp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / dp), c = (int)(i % dp);
        float hi = 0.f, lo = 0.f;
        if (c < d) {
            const double v = x[(int64_t)r * d + c];
            uint32_t hb, lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"((float)v));
            hi = __uint_as_float(hb);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"((float)(v - (double)hi)));
            lo = __uint_as_float(lb);
        }
        h[(int64_t)r * 2 * dp + c] = hi;
        h[(int64_t)r * 2 * dp + dp + c] = lo;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        HVB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { hvb_set_error("cuTensorMapEncodeTiled not available from the driver"); return HVB_ERR_UNSUPPORTED; }
        fn = (EncodeTiledFn)p;
    }
    *out = fn;
    return HVB_OK;
}

}  // namespace

int hvb_gram_tc_launch(hvb_ctx* ctx, const double* x_dev, int n, int d, float* out_g_dev) {
    const int dp = ((d + kBK - 1) / kBK) * kBK;
    float* h = nullptr;
    HVB_TRY(hvb_scratch3(ctx, (size_t)n * 2 * dp * sizeof(float), (void**)&h));
    {
        const int64_t total = (int64_t)n * dp;
        int grid = (int)((total + 255) / 256);
        if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
        split_tf32_kernel<<<grid, 256, 0, ctx->stream>>>(x_dev, n, d, dp, h);
        HVB_LAUNCHED(ctx);
    }
    EncodeTiledFn encode = nullptr;
    HVB_TRY(get_encode_fn(&encode));
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)(2 * dp), (cuuint64_t)n};
    const cuuint64_t gstride[1] = {(cuuint64_t)(2 * dp) * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)h, gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hvb_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return HVB_ERR_CUDA; }
    const size_t smem = 2 * kStages * kTileBytes + sizeof(SharedCtl) + 1024;
    HVB_CUDA(cudaFuncSetAttribute(gram_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n + kBN - 1) / kBN, (n + kBM - 1) / kBM);
    gram_tcgen05_kernel<<<grid, kThreads, smem, ctx->stream>>>(tmap, out_g_dev, n, dp);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

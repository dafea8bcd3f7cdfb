// K6 — pointwise (1x1) convolution with the K5 epilogue fused in, on the 5th-generation tensor cores.
//
// In NHWC a 1x1 convolution of the YOLO forward (ultralytics C2f.cv1 / C2f.cv2 / SPPF.cv1 / SPPF.cv2 / the last
// Conv2d of each Detect branch; reached from hockey/main.py:179-184) is a plain GEMM of two K-major operands
//     out[pixel, c_out] = act( sum_ci X[pixel, ci] * W[c_out, ci] + bias[c_out] )
// i.e. the operand layout k4_gram_tcgen05.cu already feeds to tcgen05.mma.kind::tf32, with A = activations and
// B = weights coming from two tensor maps.  cuDNN's convolution followed by the K5 bias/SiLU pass moves
// X + 3 x out through HBM; this kernel moves X + out.  Operands are the fp32 tensors themselves: the tensor core
// reads the upper 19 bits (TF32), the same precision class as the cuDNN TF32 convolutions it replaces; the
// accumulator is fp32 in TMEM.
//
// One CTA per 128-pixel x BN-channel output tile (BN = 96 or 64: every pointwise layer of YOLOv8 n/s/m has
// c_out divisible by one of them), 2-stage TMA ring over c_in in 32-float (128-byte swizzle atom) steps, three CTAs
// per SM so one tile's epilogue and start-up overlap the others' main loops (measured 10-20 % faster than a 3-stage
// ring with two CTAs per SM; HVB_PW_STAGES=3 selects that variant):
//   warp 0   TMA producer (one X box 32 x 128 and one W box 32 x BN per stage, expect_tx on an mbarrier)
//   warp 1   MMA issuer: 4 x tcgen05.mma (M128, N=BN, K8) per stage, tcgen05.commit frees the stage
//   warp 2   TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld (each warp its 32-lane quadrant) -> + bias -> activation -> 256-bit stores (128-bit when a
//            destination is only 16-byte aligned; 256-bit: 453 -> 314 us on the 96 -> 96 layer) into
//            out1[pixel * ld1 + off1 + c] and, for channels [c2_begin, c2_begin + c2_count), also into
//            out2[pixel * ld2 + off2 + c - c2_begin]  (C2f.cv1 writes the concat buffer and the dense second half)
// Rows past npix are zero-filled by TMA and never stored.  Every mbarrier wait is bounded (trap, not hang).
#include "hvb_common.cuh"

#include <cuda.h>
#include <cstdlib>

namespace {

constexpr int kBM = 128, kBK = 32;
constexpr int kMaxStages = 3;
constexpr int kXTileBytes = kBM * kBK * 4;            // 16 KB
constexpr int kThreads = 256;
constexpr unsigned kSpinLimit = 200u * 1000u * 1000u;
enum { PW_NONE = 0, PW_SILU = 1, PW_SILU_FAST = 4 };

struct SharedCtl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
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
// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (as in k4_gram_tcgen05.cu).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ float pw_silu_fast(float v) {       // same arithmetic as K5's fast SiLU
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(v, -1.4426950408889634f)));
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(1.0f, e)));
    return __fmul_rn(v, r);
}
template <int ACT>
__device__ __forceinline__ float pw_act(float v) {
    if (ACT == PW_SILU_FAST) return pw_silu_fast(v);
    if (ACT == PW_SILU) return __fdiv_rn(v, __fadd_rn(1.0f, expf(-v)));
    return v;
}

struct PwArgs {
    const float* bias;
    float* out1;
    float* out2;
    int64_t npix;
    int cin, n_tiles_n;
    int ld1, off1, ld2, off2, c2_begin, c2_count;
    int wide_stores;            // every destination row segment is 32-byte aligned: 256-bit stores
};

__device__ __forceinline__ void st256(float* p, const float4& a, const float4& b) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// STAGES = 3: two CTAs per SM; STAGES = 2 (short c_in loops): three CTAs per SM
template <int BN, int ACT, int kStages>
__global__ void __launch_bounds__(kThreads, (kStages == 2 && BN <= 128) ? 3 : 2)
pointwise_conv_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ PwArgs a) {
    constexpr int kWTileBytes = BN * kBK * 4;
    constexpr uint32_t kTmemCols = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_x = smem;
    uint8_t* smem_w = smem + kStages * kXTileBytes;
    SharedCtl* ctl = (SharedCtl*)(smem_w + kStages * kWTileBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile_m = blockIdx.x / a.n_tiles_n;
    const int tile_n = blockIdx.x - (int)(tile_m * a.n_tiles_n);
    const int64_t bm = tile_m * kBM;
    const int bn = tile_n * BN;
    const int num_kb = a.cin / kBK;

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
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                mbar_wait(&ctl->empty[s], phase ^ 1);
                mbar_expect_tx(&ctl->full[s], kXTileBytes + kWTileBytes);
                tma_load_2d(smem_x + s * kXTileBytes, &tmap_x, &ctl->full[s], kb * kBK, (int)bm);
                tma_load_2d(smem_w + s * kWTileBytes, &tmap_w, &ctl->full[s], kb * kBK, bn);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                mbar_wait(&ctl->full[s], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = make_smem_desc(smem_u32(smem_x + s * kXTileBytes));
                const uint64_t db = make_smem_desc(smem_u32(smem_w + s * kWTileBytes));
#pragma unroll
                for (int k = 0; k < kBK / 8; k++)
                    umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, (kb | k) ? 1u : 0u);
                umma_commit(&ctl->empty[s]);
            }
            umma_commit(&ctl->tmem_full);
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        mbar_wait(&ctl->tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t row = bm + q * 32 + lane;
        const bool live = row < a.npix;
        float* o1 = a.out1 + row * a.ld1 + a.off1 + bn;
        float* o2 = a.out2 ? a.out2 + row * a.ld2 + a.off2 + (bn - a.c2_begin) : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
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
            if (live) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float4 r[2];
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + bn + c + j + 4 * hh));
                        r[hh].x = pw_act<ACT>(__fadd_rn(__uint_as_float(v[j + 4 * hh]), b4.x));
                        r[hh].y = pw_act<ACT>(__fadd_rn(__uint_as_float(v[j + 4 * hh + 1]), b4.y));
                        r[hh].z = pw_act<ACT>(__fadd_rn(__uint_as_float(v[j + 4 * hh + 2]), b4.z));
                        r[hh].w = pw_act<ACT>(__fadd_rn(__uint_as_float(v[j + 4 * hh + 3]), b4.w));
                    }
                    const int ch = bn + c + j;
                    if (a.wide_stores) {             // c2_begin / c2_count are multiples of 8 in this mode
                        st256(o1 + c + j, r[0], r[1]);
                        if (o2 && ch >= a.c2_begin && ch < a.c2_begin + a.c2_count) st256(o2 + c + j, r[0], r[1]);
                    } else {
#pragma unroll
                        for (int hh = 0; hh < 2; hh++) {
                            *reinterpret_cast<float4*>(o1 + c + j + 4 * hh) = r[hh];
                            if (o2 && ch + 4 * hh >= a.c2_begin && ch + 4 * hh < a.c2_begin + a.c2_count)
                                *reinterpret_cast<float4*>(o2 + c + j + 4 * hh) = r[hh];
                        }
                    }
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int pw_encode_fn(EncodeTiledFn* out) {
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

int pw_make_map(EncodeTiledFn encode, CUtensorMap* map, const float* base, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hvb_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return HVB_ERR_CUDA; }
    return HVB_OK;
}

template <int BN, int ACT, int STAGES>
int pw_launch_s(hvb_ctx* ctx, const CUtensorMap& mx, const CUtensorMap& mw, const PwArgs& a, int64_t tiles) {
    const size_t smem = STAGES * (kXTileBytes + BN * kBK * 4) + sizeof(SharedCtl) + 1024;
    HVB_CUDA(cudaFuncSetAttribute(pointwise_conv_kernel<BN, ACT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pointwise_conv_kernel<BN, ACT, STAGES><<<(unsigned)tiles, kThreads, smem, ctx->stream>>>(mx, mw, a);
    HVB_LAUNCHED(ctx);
    return HVB_OK;
}

int pw_stages_for(int c_in) {
    static int forced = -1;                       // HVB_PW_STAGES=2|3 pins the choice (measurement knob)
    if (forced < 0) { const char* e = getenv("HVB_PW_STAGES"); forced = e ? atoi(e) : 0; }
    if (forced == 2 || forced == 3) return forced;
    (void)c_in;
    return 2;                                     // measured: three CTAs per SM beat a deeper ring at every layer size
}

template <int BN, int ACT>
int pw_launch(hvb_ctx* ctx, const CUtensorMap& mx, const CUtensorMap& mw, const PwArgs& a, int64_t tiles) {
    return pw_stages_for(a.cin) == 2 ? pw_launch_s<BN, ACT, 2>(ctx, mx, mw, a, tiles) : pw_launch_s<BN, ACT, 3>(ctx, mx, mw, a, tiles);
}

template <int BN>
int pw_dispatch(hvb_ctx* ctx, int act, const CUtensorMap& mx, const CUtensorMap& mw, const PwArgs& a, int64_t tiles) {
    switch (act) {
        case PW_NONE: return pw_launch<BN, PW_NONE>(ctx, mx, mw, a, tiles);
        case PW_SILU: return pw_launch<BN, PW_SILU>(ctx, mx, mw, a, tiles);
        default: return pw_launch<BN, PW_SILU_FAST>(ctx, mx, mw, a, tiles);
    }
}

}  // namespace

extern "C" {

int hvb_pointwise_conv(hvb_ctx* ctx, const float* x_dev, int x_ld, const float* w_dev, const float* bias_dev, int64_t npix,
                       int c_in, int c_out, int act, float* out1_dev, int out1_ld, int out1_off, float* out2_dev,
                       int out2_ld, int out2_off, int c2_begin, int c2_count) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(npix >= 0 && c_in > 0 && c_out > 0, "bad sizes");
    HVB_ARG(act == PW_NONE || act == PW_SILU || act == PW_SILU_FAST, "act must be 0 (none), 1 (SiLU) or 4 (fast SiLU)");
    if (npix == 0) return HVB_OK;
    HVB_ARG(x_dev && w_dev && bias_dev && out1_dev, "null pointer");
    if (c_in % kBK != 0 || (c_out % 96 != 0 && c_out % 64 != 0)) {
        hvb_set_error("hvb_pointwise_conv: c_in must be a multiple of 32 and c_out a multiple of 96 or 64 (got %d -> %d)", c_in, c_out);
        return HVB_ERR_UNSUPPORTED;
    }
    HVB_ARG(x_ld >= c_in && (x_ld & 3) == 0 && ((uintptr_t)x_dev & 15) == 0 && ((uintptr_t)w_dev & 15) == 0 && ((uintptr_t)bias_dev & 15) == 0,
            "x / w / bias must be 16-byte aligned and x_ld a multiple of 4 floats");
    HVB_ARG(out1_ld >= out1_off + c_out && (out1_ld & 3) == 0 && (out1_off & 3) == 0 && ((uintptr_t)out1_dev & 15) == 0, "out1 pitch / offset / alignment");
    if (out2_dev) {
        HVB_ARG(c2_begin >= 0 && c2_count > 0 && c2_begin + c2_count <= c_out && (c2_begin & 3) == 0 && (c2_count & 3) == 0, "bad out2 channel range");
        HVB_ARG(out2_ld >= out2_off + c2_count && (out2_ld & 3) == 0 && (out2_off & 3) == 0 && ((uintptr_t)out2_dev & 15) == 0, "out2 pitch / offset / alignment");
    }
    HVB_ARG(npix < ((int64_t)1 << 31), "npix too large for one tensor map");
    // (a 128 x 192 tile per CTA was measured in round 2 and removed: correct, but slower for 192 -> 192 (208 vs 187 us)
    // and only ahead for 1152 -> 384, a layer that stays on cuDNN — profiles/r02a_k6_decision.md)
    const int bn = (c_out % 96 == 0) ? 96 : 64;
    EncodeTiledFn encode = nullptr;
    HVB_TRY(pw_encode_fn(&encode));
    CUtensorMap mx, mw;
    HVB_TRY(pw_make_map(encode, &mx, x_dev, (uint64_t)c_in, (uint64_t)npix, (uint64_t)x_ld, kBM));
    HVB_TRY(pw_make_map(encode, &mw, w_dev, (uint64_t)c_in, (uint64_t)c_out, (uint64_t)c_in, (uint32_t)bn));
    PwArgs a;
    a.bias = bias_dev; a.out1 = out1_dev; a.out2 = out2_dev; a.npix = npix; a.cin = c_in; a.n_tiles_n = c_out / bn;
    a.ld1 = out1_ld; a.off1 = out1_off; a.ld2 = out2_ld; a.off2 = out2_off; a.c2_begin = c2_begin; a.c2_count = c2_count;
    static int wide_ok = -1;                      // HVB_PW_ST256=0 turns the 256-bit stores off (measurement knob)
    if (wide_ok < 0) { const char* e = getenv("HVB_PW_ST256"); wide_ok = e ? atoi(e) : 1; }
    a.wide_stores = wide_ok && (out1_ld & 7) == 0 && (out1_off & 7) == 0 && ((uintptr_t)out1_dev & 31) == 0 &&
                    (!out2_dev || ((out2_ld & 7) == 0 && (out2_off & 7) == 0 && (c2_begin & 7) == 0 && (c2_count & 7) == 0 &&
                                   ((uintptr_t)out2_dev & 31) == 0));
    const int64_t tiles = ((npix + kBM - 1) / kBM) * a.n_tiles_n;
    HVB_ARG(tiles < ((int64_t)1 << 31), "too many tiles");
    return bn == 96 ? pw_dispatch<96>(ctx, act, mx, mw, a, tiles) : pw_dispatch<64>(ctx, act, mx, mw, a, tiles);
}

}  // extern "C"

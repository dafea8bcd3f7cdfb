// Internal helpers shared by the libhvb translation units (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "hvb.h"

// Device work areas of the multi-kernel ops (K2a's counters + candidate key lists, hvb_gather_tiles' slot bases, the
// Gram operands).  One set per (stream, stream-capture id): launches that libhvb's caller issues on different streams,
// or captures into different CUDA graphs, never share an area, so they may run concurrently; launches on one stream
// are ordered by the stream.
struct hvb_work_area {
    void* k2 = nullptr;                 // [ctr | keys]
    size_t k2_bytes = 0, k2_ctr_bytes = 0;
    void* s2 = nullptr;
    size_t s2_bytes = 0;
    void* s3 = nullptr;
    size_t s3_bytes = 0;
};

struct hvb_ctx {
    int device = 0;
    int sm_count = 148;
    int max_smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    uint64_t launches = 0;
    // colour tables (uploaded once): sdiv int32[256], hdiv int32[256], gtab u16[256], ctab u16[3072]
    void* tables_dev = nullptr;
    // K3b: Pillow resampling coefficients of every source size the fast kernel accepts (built once, k3_mnv3_prep.cu)
    void* k3b_tab_dev = nullptr;
    // growable scratch used by the *_host entry points and by multi-kernel ops
    void* scratch_dev = nullptr;
    size_t scratch_bytes = 0;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    std::map<std::pair<uintptr_t, unsigned long long>, hvb_work_area> work;   // key: (stream handle, capture id or 0)
    // CUDA-graph support: once set, work buffers replaced by growth are kept alive (captured graphs may point at them)
    bool retain_buffers = false;
    std::vector<void*> retired;
};

void hvb_set_error(const char* fmt, ...);
int hvb_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int hvb_scratch(hvb_ctx* ctx, size_t bytes, void** out);     // device scratch #1 (grow-only; only the synchronous *_host entry points)
int hvb_scratch2(hvb_ctx* ctx, size_t bytes, void** out);    // device scratch #2 of the current stream's work area (grow-only)
int hvb_scratch3(hvb_ctx* ctx, size_t bytes, void** out);    // device scratch #3 of the current stream's work area (grow-only)
// K2a work area of the current stream: counters (zero between launches) + [images][cap] 64-bit keys
int hvb_k2_work(hvb_ctx* ctx, int images, int cap, unsigned long long** keys, int32_t** ctr);
int hvb_pinned(hvb_ctx* ctx, size_t bytes, void** out);      // pinned host staging (grow-only)
int hvb_k3b_build_tables(hvb_ctx* ctx);                      // k3_mnv3_prep.cu, called by hvb_ctx_create
int hvb_capturing(hvb_ctx* ctx, const char* what);           // HVB_ERR_UNSUPPORTED (with message) if ctx->stream is being captured

#define HVB_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) return hvb_cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define HVB_CHECK_CTX(ctx)                                          \
    do {                                                            \
        if (!(ctx)) { hvb_set_error("null context"); return HVB_ERR_ARG; } \
        HVB_CUDA(cudaSetDevice((ctx)->device));                     \
    } while (0)

#define HVB_ARG(cond, msg)                                  \
    do {                                                    \
        if (!(cond)) { hvb_set_error("%s: %s", __func__, msg); return HVB_ERR_ARG; } \
    } while (0)

#define HVB_TRY(expr)                       \
    do {                                    \
        int s__ = (expr);                   \
        if (s__ != HVB_OK) return s__;      \
    } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define HVB_LAUNCHED(ctx)                   \
    do {                                    \
        (ctx)->launches++;                  \
        HVB_CUDA(cudaGetLastError());       \
    } while (0)

static inline int hvb_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Byte offsets of the colour tables inside ctx->tables_dev
constexpr int HVB_TAB_SDIV = 0;                    // int32[256]
constexpr int HVB_TAB_HDIV = 1024;                 // int32[256]
constexpr int HVB_TAB_GTAB = 2048;                 // uint16[256]
constexpr int HVB_TAB_CTAB = 2560;                 // uint16[3072]
constexpr int HVB_TAB_BYTES = 2560 + 6144;         // 8704

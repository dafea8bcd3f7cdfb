// Context, error reporting and memory plumbing of libhvb (C ABI in include/hvb.h).
#include "hvb_common.cuh"

#include <algorithm>
#include <string.h>
#include <thread>

#include <math.h>
#include <string.h>

#include "hvb_tables.inc"

static thread_local char g_err[1024] = "";

void hvb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int hvb_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    (void)cudaGetLastError();   // clear the non-sticky error state so it is not re-reported by a later launch check
    hvb_set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? HVB_ERR_NO_DEVICE : HVB_ERR_CUDA;
}

int hvb_capturing(hvb_ctx* ctx, const char* what) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
        hvb_set_error("%s needs to (re)allocate a work buffer while the stream is being captured into a CUDA graph: "
                      "run the same step once eagerly before capturing", what);
        return HVB_ERR_UNSUPPORTED;
    }
    (void)cudaGetLastError();
    return HVB_OK;
}

static int grow(hvb_ctx* ctx, void** buf, size_t* have, size_t want, bool pinned) {
    if (*have >= want && *buf) return HVB_OK;
    HVB_TRY(hvb_capturing(ctx, "libhvb scratch"));
    size_t n = want + want / 4 + 4096;
    if (*buf) {
        if (ctx->retain_buffers && !pinned) {
            ctx->retired.push_back(*buf);            // a captured CUDA graph may still point at it
        } else {
            HVB_CUDA(cudaDeviceSynchronize());
            if (pinned) HVB_CUDA(cudaFreeHost(*buf)); else HVB_CUDA(cudaFree(*buf));
        }
        *buf = nullptr;
        *have = 0;
    }
    if (pinned) HVB_CUDA(cudaMallocHost(buf, n)); else HVB_CUDA(cudaMalloc(buf, n));
    *have = n;
    return HVB_OK;
}

int hvb_scratch(hvb_ctx* ctx, size_t bytes, void** out) {
    HVB_TRY(grow(ctx, &ctx->scratch_dev, &ctx->scratch_bytes, bytes, false));
    *out = ctx->scratch_dev;
    return HVB_OK;
}
// ---- per-(stream, capture) work areas -------------------------------------------------------------------------------
static int capture_id(hvb_ctx* ctx, unsigned long long* id) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    unsigned long long cid = 0;
    *id = 0;
    if (cudaStreamGetCaptureInfo(ctx->stream, &st, &cid) != cudaSuccess) { (void)cudaGetLastError(); return HVB_OK; }
    if (st == cudaStreamCaptureStatusActive) *id = cid;
    return HVB_OK;
}

static hvb_work_area* work_area(hvb_ctx* ctx, bool* capturing) {
    unsigned long long cid = 0;
    capture_id(ctx, &cid);
    *capturing = cid != 0;
    return &ctx->work[std::make_pair((uintptr_t)ctx->stream, cid)];
}

// (Re)allocate one buffer of a work area.  Outside capture the stream that owns the area is drained before the old
// buffer goes away; during capture earlier nodes of the graph may point at the old buffer, so it is retired (kept until
// the context is destroyed), and the allocation itself runs with this thread's capture mode relaxed (cudaMalloc is a
// "potentially unsafe" call under the global mode torch.cuda.graph uses).
static int work_grow(hvb_ctx* ctx, bool capturing, void** buf, size_t* have, size_t want, size_t zero_bytes) {
    if (*buf && *have >= want) return HVB_OK;
    const size_t n = want + want / 4 + 4096;
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    if (capturing) HVB_CUDA(cudaThreadExchangeStreamCaptureMode(&mode));
    cudaError_t e = cudaSuccess;
    if (*buf) {
        if (capturing || ctx->retain_buffers) {
            ctx->retired.push_back(*buf);
        } else {
            e = cudaStreamSynchronize(ctx->stream);
            if (e == cudaSuccess) e = cudaFree(*buf);
        }
        *buf = nullptr;
        *have = 0;
    }
    if (e == cudaSuccess) e = cudaMalloc(buf, n);
    if (capturing) (void)cudaThreadExchangeStreamCaptureMode(&mode);
    if (e != cudaSuccess) return hvb_cuda_fail(e, "work area allocation", __FILE__, __LINE__);
    *have = n;
    // stream-ordered before the first kernel that uses it (inside a capture: a memset node at the head of every replay)
    if (zero_bytes) HVB_CUDA(cudaMemsetAsync(*buf, 0, zero_bytes, ctx->stream));
    return HVB_OK;
}

int hvb_scratch2(hvb_ctx* ctx, size_t bytes, void** out) {
    bool cap = false;
    hvb_work_area* w = work_area(ctx, &cap);
    HVB_TRY(work_grow(ctx, cap, &w->s2, &w->s2_bytes, bytes, 0));
    *out = w->s2;
    return HVB_OK;
}
int hvb_scratch3(hvb_ctx* ctx, size_t bytes, void** out) {
    bool cap = false;
    hvb_work_area* w = work_area(ctx, &cap);
    HVB_TRY(work_grow(ctx, cap, &w->s3, &w->s3_bytes, bytes, 0));
    *out = w->s3;
    return HVB_OK;
}
int hvb_k2_work(hvb_ctx* ctx, int images, int cap, unsigned long long** keys, int32_t** ctr) {
    const size_t ctr_bytes = (((size_t)images * 2 * sizeof(int32_t)) + 255) & ~(size_t)255;
    const size_t key_bytes = (size_t)images * cap * sizeof(unsigned long long);
    bool capturing = false;
    hvb_work_area* w = work_area(ctx, &capturing);
    if (!w->k2 || ctr_bytes > w->k2_ctr_bytes || key_bytes > w->k2_bytes - w->k2_ctr_bytes) {
        const size_t ctr_cap = ctr_bytes * 2;
        size_t have = 0;                                     // force a fresh buffer: the counter / key split moves
        void* old = w->k2;
        if (old) {
            if (capturing || ctx->retain_buffers) ctx->retired.push_back(old);
            else { HVB_CUDA(cudaStreamSynchronize(ctx->stream)); HVB_CUDA(cudaFree(old)); }
            w->k2 = nullptr;
        }
        HVB_TRY(work_grow(ctx, capturing, &w->k2, &have, ctr_cap + key_bytes * 2, ctr_cap));
        w->k2_bytes = have; w->k2_ctr_bytes = ctr_cap;
    }
    *ctr = reinterpret_cast<int32_t*>(w->k2);
    *keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(w->k2) + w->k2_ctr_bytes);
    return HVB_OK;
}
int hvb_pinned(hvb_ctx* ctx, size_t bytes, void** out) {
    HVB_TRY(grow(ctx, &ctx->pinned, &ctx->pinned_bytes, bytes, true));
    *out = ctx->pinned;
    return HVB_OK;
}

extern "C" {

int hvb_version(void) { return HVB_VERSION; }

const char* hvb_last_error(void) { return g_err; }

int hvb_device_count(int* out_count) {
    if (!out_count) { hvb_set_error("null out_count"); return HVB_ERR_ARG; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    *out_count = n;
    return HVB_OK;
}

int hvb_ctx_create(int device, hvb_ctx** out_ctx) {
    if (!out_ctx) { hvb_set_error("null out_ctx"); return HVB_ERR_ARG; }
    *out_ctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        hvb_set_error("no CUDA device available (%s); libhvb has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return HVB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { hvb_set_error("device %d out of range [0,%d)", device, n); return HVB_ERR_ARG; }
    HVB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HVB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        hvb_set_error("device %d is sm_%d%d; libhvb is built for sm_100a only", device, prop.major, prop.minor);
        return HVB_ERR_NO_DEVICE;
    }
    hvb_ctx* c = new hvb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    HVB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    HVB_CUDA(cudaEventCreate(&c->ev_start));
    HVB_CUDA(cudaEventCreate(&c->ev_stop));

    // colour tables: HSV division tables (exact in double) + the generated LAB tables
    std::vector<unsigned char> tab(HVB_TAB_BYTES);
    int32_t* sdiv = (int32_t*)(tab.data() + HVB_TAB_SDIV);
    int32_t* hdiv = (int32_t*)(tab.data() + HVB_TAB_HDIV);
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; i++) {
        sdiv[i] = (int32_t)nearbyint((double)(255 << 12) / (double)i);
        hdiv[i] = (int32_t)nearbyint((double)(180 << 12) / (6.0 * (double)i));
    }
    memcpy(tab.data() + HVB_TAB_GTAB, kHvbLabGammaTab, sizeof(kHvbLabGammaTab));
    memcpy(tab.data() + HVB_TAB_CTAB, kHvbLabCbrtTab, sizeof(kHvbLabCbrtTab));
    HVB_CUDA(cudaMalloc(&c->tables_dev, HVB_TAB_BYTES));
    HVB_CUDA(cudaMemcpy(c->tables_dev, tab.data(), HVB_TAB_BYTES, cudaMemcpyHostToDevice));
    {
        int st = hvb_k3b_build_tables(c);
        if (st != HVB_OK) return st;
    }
    // the copy above is ordered on the legacy stream only; the kernels run on non-blocking streams that do not
    // synchronise with it, so make the upload globally visible before the context is handed out
    HVB_CUDA(cudaDeviceSynchronize());
    *out_ctx = c;
    return HVB_OK;
}

int hvb_ctx_retain_buffers(hvb_ctx* ctx, int on) {
    if (!ctx) { hvb_set_error("null context"); return HVB_ERR_ARG; }
    ctx->retain_buffers = on != 0;
    return HVB_OK;
}

int hvb_ctx_destroy(hvb_ctx* ctx) {
    if (!ctx) return HVB_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->tables_dev) cudaFree(ctx->tables_dev);
    if (ctx->k3b_tab_dev) cudaFree(ctx->k3b_tab_dev);
    if (ctx->scratch_dev) cudaFree(ctx->scratch_dev);
    for (auto& kv : ctx->work) {
        if (kv.second.k2) cudaFree(kv.second.k2);
        if (kv.second.s2) cudaFree(kv.second.s2);
        if (kv.second.s3) cudaFree(kv.second.s3);
    }
    for (void* p : ctx->retired) cudaFree(p);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return HVB_OK;
}

int hvb_ctx_set_stream(hvb_ctx* ctx, void* cuda_stream) {
    if (!ctx) { hvb_set_error("null context"); return HVB_ERR_ARG; }
    ctx->stream = (cudaStream_t)cuda_stream;      // NULL is the legacy default stream (what torch uses by default)
    return HVB_OK;
}

int hvb_ctx_use_own_stream(hvb_ctx* ctx) {
    if (!ctx) { hvb_set_error("null context"); return HVB_ERR_ARG; }
    ctx->stream = ctx->own_stream;
    return HVB_OK;
}

int hvb_ctx_get_stream(hvb_ctx* ctx, void** out_stream) {
    if (!ctx || !out_stream) { hvb_set_error("null argument"); return HVB_ERR_ARG; }
    *out_stream = (void*)ctx->stream;
    return HVB_OK;
}

int hvb_ctx_synchronize(hvb_ctx* ctx) {
    HVB_CHECK_CTX(ctx);
    HVB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HVB_OK;
}

int hvb_ctx_sm_count(hvb_ctx* ctx, int* out_sms) {
    if (!ctx || !out_sms) { hvb_set_error("null argument"); return HVB_ERR_ARG; }
    *out_sms = ctx->sm_count;
    return HVB_OK;
}

int hvb_malloc(hvb_ctx* ctx, size_t bytes, void** out_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(out_dev != nullptr, "null out_dev");
    HVB_CUDA(cudaMalloc(out_dev, bytes ? bytes : 1));
    return HVB_OK;
}

int hvb_free(hvb_ctx* ctx, void* ptr_dev) {
    HVB_CHECK_CTX(ctx);
    if (ptr_dev) HVB_CUDA(cudaFree(ptr_dev));
    return HVB_OK;
}

int hvb_host_alloc(hvb_ctx* ctx, size_t bytes, void** out_host) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(out_host != nullptr, "null out_host");
    HVB_CUDA(cudaMallocHost(out_host, bytes ? bytes : 1));
    return HVB_OK;
}

int hvb_host_free(hvb_ctx* ctx, void* ptr_host) {
    HVB_CHECK_CTX(ctx);
    if (ptr_host) HVB_CUDA(cudaFreeHost(ptr_host));
    return HVB_OK;
}

int hvb_memcpy_h2d(hvb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    HVB_CHECK_CTX(ctx);
    if (bytes) HVB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return HVB_OK;
}

// Host-side staging of a chunk of decoded frames (hockey/main.py:321: every frame is a separate host array) into ONE
// contiguous — normally pinned — buffer, the source of the chunk's single H2D copy.  n_threads workers copy disjoint
// frames; measured on the gpurun box, a single-threaded 398 MB staging copy (64 x 1080p) took as long as the whole GPU
// step and was the ceiling of the end-to-end figure.  No CUDA call, no context.
int hvb_stage_frames(const void* const* src_host, int n, size_t bytes_each, void* dst_host, int n_threads) {
    HVB_ARG(n >= 0 && (n == 0 || (src_host && dst_host)), "null pointer");
    if (n == 0 || bytes_each == 0) return HVB_OK;
    for (int k = 0; k < n; k++) HVB_ARG(src_host[k] != nullptr, "null frame pointer");
    const int t = std::max(1, std::min(n_threads, std::min(n, 64)));
    auto work = [&](int lane) {
        for (int k = lane; k < n; k += t) memcpy((uint8_t*)dst_host + (size_t)k * bytes_each, src_host[k], bytes_each);
    };
    if (t == 1) { work(0); return HVB_OK; }
    std::vector<std::thread> pool;
    pool.reserve(t - 1);
    for (int lane = 1; lane < t; lane++) pool.emplace_back(work, lane);
    work(0);
    for (auto& th : pool) th.join();
    return HVB_OK;
}

int hvb_memcpy_d2h(hvb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    HVB_CHECK_CTX(ctx);
    if (bytes) HVB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return HVB_OK;
}

int hvb_memset(hvb_ctx* ctx, void* dst_dev, int value, size_t bytes) {
    HVB_CHECK_CTX(ctx);
    if (bytes) HVB_CUDA(cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
    return HVB_OK;
}

int hvb_ctx_launch_count(hvb_ctx* ctx, int reset, uint64_t* out_launches) {
    if (!ctx) { hvb_set_error("null context"); return HVB_ERR_ARG; }
    if (out_launches) *out_launches = ctx->launches;
    if (reset) ctx->launches = 0;
    return HVB_OK;
}

int hvb_timer_start(hvb_ctx* ctx) {
    HVB_CHECK_CTX(ctx);
    HVB_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
    return HVB_OK;
}

int hvb_timer_stop_ms(hvb_ctx* ctx, float* out_ms) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(out_ms != nullptr, "null out_ms");
    HVB_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
    HVB_CUDA(cudaEventSynchronize(ctx->ev_stop));
    HVB_CUDA(cudaEventElapsedTime(out_ms, ctx->ev_start, ctx->ev_stop));
    return HVB_OK;
}

}  // extern "C"

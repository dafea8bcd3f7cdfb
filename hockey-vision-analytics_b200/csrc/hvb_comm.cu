// Feature exchange for the global team-clustering fit (SURVEY.md §8b `hvb_allgather_features`, §8e): the ONE collective
// on the hot path.  Each rank holds the crop features of its clips, float64[n_g, 625]; before
// HybridTeamClassifier.fit (hockey/common/team_hybrid.py:155-196) standardises them, every rank needs the rows of all
// ranks, in rank order, bit-identical (so that the scaler statistics are identical everywhere).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): libhvb has no link-time dependency on it, a single-GPU host never
// loads it, and a host that already carries NCCL (PyTorch does) shares its copy.  The exchange itself is an
// all-gather-v expressed as one ncclGroup of per-rank broadcasts straight into the compacted output — no padding to the
// largest rank and no compaction pass: rank r's rows land at row offset sum(counts[:r]).  Messages are <= ~10 MB
// (latency-bound over NVLink 5 / NVSwitch); there is no compute step to fuse with: the Gram needs the standardised
// rows, and the scaler needs every row first.
#include "hvb_common.cuh"

#include <dlfcn.h>
#include <mutex>
#include <string.h>

namespace {

// The handful of NCCL declarations this file needs (ABI-stable across NCCL 2.x: nccl.h of 2.27 / 2.28).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;            // ncclSuccess == 0
enum { kNcclInt32 = 2, kNcclFloat64 = 8 };       // ncclInt32 / ncclFloat64 in ncclDataType_t

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("HVB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = "libnccl.so.2 not found (set HVB_NCCL_LIB)"; return; }
        struct { const char* name; void** fn; } syms[] = {
            {"ncclGetUniqueId", (void**)&api.GetUniqueId}, {"ncclCommInitRank", (void**)&api.CommInitRank},
            {"ncclCommDestroy", (void**)&api.CommDestroy}, {"ncclCommCount", (void**)&api.CommCount},
            {"ncclCommUserRank", (void**)&api.CommUserRank}, {"ncclAllGather", (void**)&api.AllGather},
            {"ncclBroadcast", (void**)&api.Broadcast}, {"ncclGroupStart", (void**)&api.GroupStart},
            {"ncclGroupEnd", (void**)&api.GroupEnd}, {"ncclGetErrorString", (void**)&api.GetErrorString}};
        for (auto& s : syms) {
            *s.fn = dlsym(api.handle, s.name);
            if (!*s.fn) { api.error = std::string("NCCL symbol missing: ") + s.name; api.handle = nullptr; return; }
        }
    });
    return &api;
}

int nccl_fail(NcclApi* api, ncclResult_t r, const char* what) {
    hvb_set_error("NCCL: %s failed: %s", what, api->GetErrorString ? api->GetErrorString(r) : "?");
    return HVB_ERR_CUDA;
}

#define HVB_NCCL(api, call)                                   \
    do {                                                      \
        ncclResult_t r__ = (api)->call;                       \
        if (r__ != 0) return nccl_fail((api), r__, #call);    \
    } while (0)

}  // namespace

struct hvb_comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    bool owned = false;
    int32_t* counts_dev = nullptr;       // [world]
    int32_t* counts_pinned = nullptr;    // [world + 1]: counts + this rank's own count as the send buffer source
};

extern "C" {

int hvb_comm_unique_id(uint8_t* out_id128) {
    HVB_ARG(out_id128 != nullptr, "null id buffer");
    NcclApi* api = nccl_api();
    if (!api->handle) { hvb_set_error("hvb_comm_unique_id: %s", api->error.c_str()); return HVB_ERR_UNSUPPORTED; }
    ncclUniqueId id;
    HVB_NCCL(api, GetUniqueId(&id));
    memcpy(out_id128, id.internal, HVB_COMM_ID_BYTES);
    return HVB_OK;
}

static int comm_finish(hvb_comm* c, hvb_comm** out) {
    cudaError_t e = cudaMalloc(&c->counts_dev, sizeof(int32_t) * (size_t)(c->world + 1));
    if (e == cudaSuccess) e = cudaMallocHost(&c->counts_pinned, sizeof(int32_t) * (size_t)(c->world + 1));
    if (e != cudaSuccess) {
        if (c->counts_dev) cudaFree(c->counts_dev);
        delete c;
        return hvb_cuda_fail(e, "hvb_comm: buffers", __FILE__, __LINE__);
    }
    *out = c;
    return HVB_OK;
}

int hvb_comm_create(hvb_ctx* ctx, const uint8_t* id128, int world, int rank, hvb_comm** out_comm) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(id128 && out_comm, "null pointer");
    HVB_ARG(world >= 1 && rank >= 0 && rank < world, "bad world / rank");
    *out_comm = nullptr;
    NcclApi* api = nccl_api();
    if (!api->handle) { hvb_set_error("hvb_comm_create: %s", api->error.c_str()); return HVB_ERR_UNSUPPORTED; }
    ncclUniqueId id;
    memcpy(id.internal, id128, HVB_COMM_ID_BYTES);
    hvb_comm* c = new hvb_comm();
    c->world = world; c->rank = rank; c->owned = true;
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != 0) { delete c; return nccl_fail(api, r, "ncclCommInitRank"); }
    return comm_finish(c, out_comm);
}

int hvb_comm_wrap(hvb_ctx* ctx, void* nccl_comm, hvb_comm** out_comm) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(nccl_comm && out_comm, "null pointer");
    *out_comm = nullptr;
    NcclApi* api = nccl_api();
    if (!api->handle) { hvb_set_error("hvb_comm_wrap: %s", api->error.c_str()); return HVB_ERR_UNSUPPORTED; }
    hvb_comm* c = new hvb_comm();
    c->comm = (ncclComm_t)nccl_comm;
    ncclResult_t r = api->CommCount(c->comm, &c->world);
    if (r == 0) r = api->CommUserRank(c->comm, &c->rank);
    if (r != 0) { delete c; return nccl_fail(api, r, "ncclCommCount / ncclCommUserRank"); }
    return comm_finish(c, out_comm);
}

int hvb_comm_destroy(hvb_ctx* ctx, hvb_comm* comm) {
    if (!comm) return HVB_OK;
    HVB_CHECK_CTX(ctx);
    HVB_CUDA(cudaStreamSynchronize(ctx->stream));
    NcclApi* api = nccl_api();
    if (comm->owned && comm->comm && api->handle) api->CommDestroy(comm->comm);
    if (comm->counts_dev) cudaFree(comm->counts_dev);
    if (comm->counts_pinned) cudaFreeHost(comm->counts_pinned);
    delete comm;
    return HVB_OK;
}

int hvb_comm_info(hvb_comm* comm, int* out_world, int* out_rank) {
    HVB_ARG(comm != nullptr, "null communicator");
    if (out_world) *out_world = comm->world;
    if (out_rank) *out_rank = comm->rank;
    return HVB_OK;
}

// Phase 1: exchange the row counts (synchronises the context's stream: the host sizes the output from them).
int hvb_allgather_counts(hvb_ctx* ctx, hvb_comm* comm, int n_local, int32_t* out_counts_host, int64_t* out_total) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(comm && out_counts_host, "null pointer");
    HVB_ARG(n_local >= 0, "negative row count");
    HVB_TRY(hvb_capturing(ctx, "hvb_allgather_counts"));
    NcclApi* api = nccl_api();
    const int w = comm->world;
    comm->counts_pinned[w] = n_local;
    HVB_CUDA(cudaMemcpyAsync(comm->counts_dev + w, comm->counts_pinned + w, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    HVB_NCCL(api, AllGather(comm->counts_dev + w, comm->counts_dev, 1, kNcclInt32, comm->comm, ctx->stream));
    HVB_CUDA(cudaMemcpyAsync(comm->counts_pinned, comm->counts_dev, sizeof(int32_t) * (size_t)w, cudaMemcpyDeviceToHost, ctx->stream));
    HVB_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t total = 0;
    for (int r = 0; r < w; r++) { out_counts_host[r] = comm->counts_pinned[r]; total += comm->counts_pinned[r]; }
    if (out_total) *out_total = total;
    return HVB_OK;
}

// Phase 2: the rows.  counts_host: what phase 1 returned; out_dev: float64[sum counts, d], rank order.  Asynchronous on
// the context's stream.
int hvb_allgather_features(hvb_ctx* ctx, hvb_comm* comm, const double* local_dev, int n_local, int d,
                           const int32_t* counts_host, double* out_dev) {
    HVB_CHECK_CTX(ctx);
    HVB_ARG(comm && counts_host && out_dev, "null pointer");
    HVB_ARG(d >= 1 && n_local >= 0 && (n_local == 0 || local_dev), "bad sizes");
    HVB_ARG(counts_host[comm->rank] == n_local, "counts_host[rank] differs from n_local");
    NcclApi* api = nccl_api();
    int64_t off = 0;
    HVB_NCCL(api, GroupStart());
    for (int r = 0; r < comm->world; r++) {
        const size_t elems = (size_t)counts_host[r] * (size_t)d;
        if (elems) {
            ncclResult_t rr = api->Broadcast(r == comm->rank ? (const void*)local_dev : (const void*)(out_dev + off), out_dev + off,
                                             elems, kNcclFloat64, r, comm->comm, ctx->stream);
            if (rr != 0) { api->GroupEnd(); return nccl_fail(api, rr, "ncclBroadcast"); }
        }
        off += (int64_t)elems;
    }
    HVB_NCCL(api, GroupEnd());
    return HVB_OK;
}

}  // extern "C"

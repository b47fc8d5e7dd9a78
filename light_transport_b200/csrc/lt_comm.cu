// lt_comm.cu — the one collective of the path behind the C-ABI: the per-rank films are summed by a single NCCL reduce
// over NVLink / NVSwitch at the end of a pass, fused with the finalize on the root.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy PyTorch has already loaded when the host is a
// torch.distributed job, else the system's), so liblt_b200.so still loads where NCCL is absent; these entry points
// then fail with LT_E_NCCL.  The few declarations needed are restated here (nccl.h 2.27: ncclUniqueId is 128 opaque
// bytes passed by value; ncclFloat64 = 8, ncclSum = 0, ncclSuccess = 0, ncclInProgress = 7).
// Failures do not hang the caller: every call is followed by ncclCommGetAsyncError until the work is enqueued, and
// an asynchronous error aborts the communicator and comes back as a status code.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

#include "lt_common.cuh"
#include "lt_host.h"

namespace lt {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclInProgress = 7, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*CommAbort)(ncclComm_t) = nullptr;
    int (*CommGetAsyncError)(ncclComm_t, int*) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static const NcclApi& nccl() {
    std::call_once(g_nccl_once, [] {
        NcclApi& a = g_nccl;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return;
#define LT_SYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym))
        LT_SYM(GetUniqueId, "ncclGetUniqueId"); LT_SYM(CommInitRank, "ncclCommInitRank"); LT_SYM(CommDestroy, "ncclCommDestroy");
        LT_SYM(CommAbort, "ncclCommAbort"); LT_SYM(CommGetAsyncError, "ncclCommGetAsyncError"); LT_SYM(Reduce, "ncclReduce");
        LT_SYM(AllReduce, "ncclAllReduce"); LT_SYM(GetErrorString, "ncclGetErrorString"); LT_SYM(GetVersion, "ncclGetVersion");
#undef LT_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.CommAbort && a.CommGetAsyncError && a.Reduce && a.AllReduce && a.GetErrorString;
    });
    return g_nccl;
}

__global__ void finalize_div_inplace_kernel(double* film, int64_t n, double spp) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = film[i] / spp;       // R/src/path_tracing.py:304: color/scene.number_of_samples, then np.clip(color, 0, 1)
        film[i] = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    }
}

}  // namespace lt

using namespace lt;

struct lt_comm {
    lt_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int n_ranks = 0, rank = 0;
    bool broken = false;
};

static int need_nccl(const char* who) {
    if (nccl().ok) return LT_OK;
    set_error("%s: NCCL is not available (dlopen of libnccl.so.2 failed or symbols are missing)", who);
    return LT_E_NCCL;
}

// a failed or asynchronously failed call: abort the communicator (never hang in a later collective) and report
static int nccl_fail(lt_comm* c, const char* what, int rc) {
    set_error("%s failed: %s (NCCL result %d, rank %d of %d)", what, nccl().GetErrorString(rc), rc, c ? c->rank : -1, c ? c->n_ranks : 0);
    if (c && c->comm) { nccl().CommAbort(c->comm); c->comm = nullptr; c->broken = true; }
    return LT_E_NCCL;
}

// after enqueuing: poll the communicator's asynchronous state until it is no longer "in progress"
static int nccl_settle(lt_comm* c, const char* what) {
    for (int spins = 0;; ++spins) {
        int st = kNcclSuccess;
        const int rc = nccl().CommGetAsyncError(c->comm, &st);
        if (rc != kNcclSuccess) return nccl_fail(c, what, rc);
        if (st == kNcclSuccess) return LT_OK;
        if (st != kNcclInProgress) return nccl_fail(c, what, st);
        if (spins > 600000) return nccl_fail(c, what, st);                     // ~60 s of "in progress": give up rather than hang
        std::this_thread::sleep_for(std::chrono::microseconds(100));
    }
}

extern "C" {

int lt_comm_unique_id(void* id128) {
    if (!id128) { set_error("lt_comm_unique_id: null argument"); return LT_E_INVALID; }
    int rc = need_nccl("lt_comm_unique_id");
    if (rc) return rc;
    ncclUniqueId id;
    const int r = nccl().GetUniqueId(&id);
    if (r != kNcclSuccess) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    memcpy(id128, id.internal, sizeof id.internal);
    return LT_OK;
}

int lt_comm_init(lt_ctx* ctx, int n_ranks, int rank, const void* id128, lt_comm** out) {
    if (!ctx || !id128 || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("lt_comm_init: bad argument"); return LT_E_INVALID; }
    *out = nullptr;
    int rc = need_nccl("lt_comm_init");
    if (rc) return rc;
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    lt_comm* c = new lt_comm();
    c->ctx = ctx; c->n_ranks = n_ranks; c->rank = rank;
    ncclUniqueId id;
    memcpy(id.internal, id128, sizeof id.internal);
    const int r = nccl().CommInitRank(&c->comm, n_ranks, id, rank);
    if (r != kNcclSuccess) { rc = nccl_fail(c, "ncclCommInitRank", r); delete c; return rc; }
    if ((rc = nccl_settle(c, "ncclCommInitRank"))) { delete c; return rc; }
    *out = c;
    return LT_OK;
}

void lt_comm_destroy(lt_comm* c) {
    if (!c) return;
    if (c->comm && nccl().ok) {
        cudaSetDevice(c->ctx->device);
        nccl().CommDestroy(c->comm);
    }
    delete c;
}

int lt_film_reduce(lt_ctx* ctx, lt_comm* c, double* film_sum, int64_t n, int root, int64_t spp_total) {
    if (!ctx || !c || !film_sum || n < 0 || root >= c->n_ranks || c->ctx != ctx) { set_error("lt_film_reduce: bad argument"); return LT_E_INVALID; }
    if (c->broken || !c->comm) { set_error("lt_film_reduce: the communicator was aborted by an earlier NCCL failure"); return LT_E_NCCL; }
    int rc = need_nccl("lt_film_reduce");
    if (rc) return rc;
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, film_sum) != cudaSuccess || (pa.type != cudaMemoryTypeDevice && pa.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        set_error("lt_film_reduce: film_sum must be device memory");
        return LT_E_INVALID;
    }
    cudaStream_t st = ctx->stream;
    if (n > 0 && c->n_ranks > 1) {
        const int r = root < 0 ? nccl().AllReduce(film_sum, film_sum, (size_t)n, kNcclFloat64, kNcclSum, c->comm, st)
                               : nccl().Reduce(film_sum, film_sum, (size_t)n, kNcclFloat64, kNcclSum, root, c->comm, st);
        if (r != kNcclSuccess && r != kNcclInProgress) return nccl_fail(c, root < 0 ? "ncclAllReduce" : "ncclReduce", r);
        if ((rc = nccl_settle(c, root < 0 ? "ncclAllReduce" : "ncclReduce"))) return rc;
    }
    // fused finalize where the sum lands: image = clip(film / spp, 0, 1), in place, same stream
    if (spp_total > 0 && n > 0 && (root < 0 || root == c->rank)) {
        const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->prop.multiProcessorCount * 8);
        finalize_div_inplace_kernel<<<std::max(grid, 1), 256, 0, st>>>(film_sum, n, (double)spp_total);
        LT_CUDA_OK(cudaGetLastError());
    }
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        // a failed collective surfaces as a stream error: say which and drop the communicator
        int ast = kNcclSuccess;
        if (c->comm) nccl().CommGetAsyncError(c->comm, &ast);
        set_error("lt_film_reduce: stream failed: %s (NCCL async state: %s)", cudaGetErrorString(e), nccl().GetErrorString(ast));
        if (c->comm) { nccl().CommAbort(c->comm); c->comm = nullptr; c->broken = true; }
        return ast != kNcclSuccess ? LT_E_NCCL : LT_E_CUDA;
    }
    return nccl_settle(c, "lt_film_reduce");
}

int lt_nccl_version(int* version) {
    if (!version) { set_error("null argument"); return LT_E_INVALID; }
    int rc = need_nccl("lt_nccl_version");
    if (rc) return rc;
    if (!nccl().GetVersion) { *version = 0; return LT_OK; }
    return nccl().GetVersion(version) == kNcclSuccess ? LT_OK : LT_E_NCCL;
}

}  // extern "C"

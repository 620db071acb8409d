// comm.cu — the exchange steps of the path: one allreduce of a d-vector (+ a scalar) per row-sharded streaming pass
//  and, when the pass feeds a replicated inner epoch, one all-gather of the per-row step scalars.  One process per GPU;
// NCCL is bound at run time with dlopen so that libciao_cuda has no link-time
// dependency on it and shares the copy already loaded by the host process.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace {
struct NcclUid { char internal[128]; };
typedef int (*fn_GetUniqueId)(NcclUid *);
typedef int (*fn_CommInitRank)(void **, int, NcclUid, int);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_AllGather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*fn_CommDestroy)(void *);
typedef const char *(*fn_GetErrorString)(int);
struct NcclApi {
    void *h = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
} g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

int nccl_load() {
    if (g_nccl.h) return CIAO_OK;
    const char *cands[] = {getenv("CIAO_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *name : cands) {
        if (!name) continue;
        g_nccl.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) CIAO_FAIL(CIAO_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_CommInitRank)dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.AllReduce = (fn_AllReduce)dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.AllGather = (fn_AllGather)dlsym(g_nccl.h, "ncclAllGather");
    g_nccl.CommDestroy = (fn_CommDestroy)dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.GetErrorString = (fn_GetErrorString)dlsym(g_nccl.h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        g_nccl.h = nullptr;
        CIAO_FAIL(CIAO_ERR_COMM, "libnccl lacks a required symbol");
    }
    return CIAO_OK;
}
#define NCCL_TRY(expr)                                                                                     \
    do {                                                                                                   \
        int _r = (expr);                                                                                   \
        if (_r != 0)                                                                                       \
            CIAO_FAIL(CIAO_ERR_COMM, "%s -> %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error"); \
    } while (0)
}  // namespace

void ciao_comm_destroy(ciao_ctx *c);

extern "C" int ciao_comm_unique_id(void *out128) {
    if (!out128) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_comm_unique_id: null output");
    CIAO_TRY(nccl_load());
    NcclUid id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return CIAO_OK;
}

extern "C" int ciao_comm_init(ciao_ctx *c, const void *id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_comm_init: bad arguments");
    ciao_comm_destroy(c);   // a second ciao_comm_init replaces the communicator instead of leaking it
    if (world == 1) {
        c->rank = 0;
        c->world = 1;
        return CIAO_OK;
    }
    CIAO_TRY(nccl_load());
    CUDA_TRY(cudaSetDevice(c->device));
    NcclUid id;
    memcpy(&id, id128, sizeof(id));
    NCCL_TRY(g_nccl.CommInitRank(&c->nccl_comm, world, id, rank));
    c->rank = rank;
    c->world = world;
    return CIAO_OK;
}

// in-place allreduce of `count` doubles on the context stream (sum, or max)
int ciao_comm_allreduce(ciao_ctx *c, double *buf, int64_t count, int op_max) {
    if (c->world <= 1) return CIAO_OK;
    if (!c->nccl_comm) CIAO_FAIL(CIAO_ERR_STATE, "allreduce: communicator not initialised");
    NCCL_TRY(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat64, op_max ? kNcclMax : kNcclSum, c->nccl_comm, c->stream));
    return CIAO_OK;
}

// in-place all-gather: rank r contributes buf[r·count .. (r+1)·count), every rank ends with all world·count doubles.
// The second exchange of a row-sharded full-gradient pass: the per-row step scalars {b_i, λ_i, 0, c_i(z_full)} each rank
// computed for its window, needed by the (replicated) sequential inner epoch.
int ciao_comm_allgather_inplace(ciao_ctx *c, double *buf, int64_t count_per_rank) {
    if (c->world <= 1) return CIAO_OK;
    if (!c->nccl_comm || !g_nccl.AllGather) CIAO_FAIL(CIAO_ERR_STATE, "allgather: communicator not initialised");
    NCCL_TRY(g_nccl.AllGather(buf + (size_t)c->rank * count_per_rank, buf, (size_t)count_per_rank, kNcclFloat64, c->nccl_comm, c->stream));
    return CIAO_OK;
}

void ciao_comm_destroy(ciao_ctx *c) {
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    c->nccl_comm = nullptr;
}

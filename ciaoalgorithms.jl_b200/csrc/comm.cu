// comm.cu — the exchange steps of the path: one allreduce of a d-vector (+ a scalar) per row-sharded streaming pass
//  and, when the pass feeds a replicated inner epoch, one all-gather of the per-row step scalars.  One process per GPU;
// NCCL is bound at run time with dlopen so that libciao_cuda has no link-time
// dependency on it and shares the copy already loaded by the host process.
#include <dlfcn.h>
#include <string.h>
#include <unistd.h>

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
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);   // a second ciao_comm_init replaces the communicator
    c->nccl_comm = nullptr;
    if (c->p2p_ready && (c->rank != rank || c->world != world))
        CIAO_FAIL(CIAO_ERR_INVALID, "ciao_comm_init: rank/world differ from the attached peer exchange's (%d/%d)", c->rank, c->world);
    if (world == 1) {
        c->rank = 0;
        c->world = 1;
        c->p2p_ready = false;
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

// ---------------------------------------------------------------------------
// One-shot exchange over peer memory (SURVEY.md §5 last row, §8e / App. B.5): the deterministic replacement of the
// NCCL allreduce on the pass path.  Every rank owns an arena (mail slots + flags, common.cuh) that its peers map — CUDA IPC
// between processes, plain device pointers (+ cudaDeviceEnablePeerAccess) inside one process — and the tail kernel of a pass
// (pass.cu pass_tail_kernel) stores its partial d-vector into every peer's slot, raises a flag, waits for the peers' flags
// and sums the slots in rank order: the same bits on every rank, run to run, with one kernel after the pass instead of
// reduce + ncclAllReduce + finish.  Works between processes that share one GPU too (the driver's 1-GPU test box).
// ---------------------------------------------------------------------------
namespace {
struct P2PHandle {             // 128 bytes, opaque to the caller
    uint64_t magic, pid, ptr;
    int32_t device, pad;
    cudaIpcMemHandle_t ipc;    // 64 bytes
    char fill[128 - 32 - 64];
};
static_assert(sizeof(P2PHandle) == 128, "p2p handle blob is 128 bytes");
constexpr uint64_t kP2PMagic = 0x4349414f50325031ull;  // "CIAOP2P1"
}  // namespace

static void p2p_detach(ciao_ctx *c) {
    for (int r = 0; r < CIAO_MAX_PEERS; ++r) {
        if (c->p2p_mapped[r]) cudaIpcCloseMemHandle(c->p2p_mapped[r]);
        c->p2p_mapped[r] = nullptr;
        c->p2p_peer[r] = nullptr;
    }
    c->p2p_ready = false;
}

extern "C" int ciao_comm_p2p_handle(ciao_ctx *c, void *out128) {
    if (!c || !out128) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_comm_p2p_handle: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (!c->p2p_arena) {
        CUDA_TRY(cudaMalloc(&c->p2p_arena, P2P_ARENA_BYTES));
        CUDA_TRY(cudaMemset(c->p2p_arena, 0, P2P_ARENA_BYTES));   // flags start at 0 before anybody can hold the handle
        CUDA_TRY(cudaDeviceSynchronize());
    }
    P2PHandle h;
    memset(&h, 0, sizeof(h));
    h.magic = kP2PMagic;
    h.pid = (uint64_t)getpid();
    h.ptr = (uint64_t)(uintptr_t)c->p2p_arena;
    h.device = c->device;
    CUDA_TRY(cudaIpcGetMemHandle(&h.ipc, c->p2p_arena));
    memcpy(out128, &h, sizeof(h));
    return CIAO_OK;
}

extern "C" int ciao_comm_p2p_attach(ciao_ctx *c, int rank, int world, const void *handles128) {
    if (!c || !handles128 || world < 1 || world > CIAO_MAX_PEERS || rank < 0 || rank >= world)
        CIAO_FAIL(CIAO_ERR_INVALID, "ciao_comm_p2p_attach: 1..%d ranks, rank inside", CIAO_MAX_PEERS);
    if (!c->p2p_arena) CIAO_FAIL(CIAO_ERR_STATE, "ciao_comm_p2p_attach before ciao_comm_p2p_handle");
    if (c->nccl_comm && (c->rank != rank || c->world != world))
        CIAO_FAIL(CIAO_ERR_INVALID, "ciao_comm_p2p_attach: rank/world differ from the NCCL communicator's (%d/%d)", c->rank, c->world);
    CUDA_TRY(cudaSetDevice(c->device));
    p2p_detach(c);
    const P2PHandle *hs = (const P2PHandle *)handles128;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            c->p2p_peer[r] = c->p2p_arena;
            continue;
        }
        P2PHandle h;
        memcpy(&h, hs + r, sizeof(h));
        if (h.magic != kP2PMagic) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_comm_p2p_attach: entry %d is not a ciao_comm_p2p_handle blob", r);
        if (h.pid == (uint64_t)getpid()) {   // same process: the pointer is valid as is; another device needs peer access
            if (h.device != c->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    ciao_set_error("cudaDeviceEnablePeerAccess(%d) from %d: %s", h.device, c->device, cudaGetErrorString(e));
                    return CIAO_ERR_COMM;
                }
                cudaGetLastError();
            }
            c->p2p_peer[r] = (double *)(uintptr_t)h.ptr;
        } else {
            void *ptr = nullptr;
            CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            c->p2p_mapped[r] = ptr;
            c->p2p_peer[r] = (double *)ptr;
        }
    }
    if (const char *t = getenv("CIAO_P2P_TIMEOUT_MS")) c->p2p_timeout_ns = (uint64_t)atoll(t) * 1000000ull;
    c->rank = rank;
    c->world = world;
    c->p2p_ready = world > 1;
    return CIAO_OK;
}

int run_p2p_allreduce(ciao_ctx *c, double *buf, int64_t count, int op_max);   // pass.cu (tail kernel on a plain buffer)

// in-place allreduce of `count` doubles on the context stream (sum, or max)
int ciao_comm_allreduce(ciao_ctx *c, double *buf, int64_t count, int op_max) {
    if (c->world <= 1) return CIAO_OK;
    if (c->p2p_ready) return run_p2p_allreduce(c, buf, count, op_max);
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
    p2p_detach(c);
    if (c->p2p_arena) cudaFree(c->p2p_arena);
    c->p2p_arena = nullptr;
    c->p2p_seq = 0;
    c->p2p_ll_epoch = 0;
}

// abi.cu — the C ABI of libciao_cuda (include/ciao_cuda.h): context lifetime,
// problem upload, the solver entry points that replace Base.iterate of the
// reference's iterables, state read-back and measurement.  Unity build: the
// kernel files are included here so that every kernel lives in one module.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";
void ciao_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#include "pass.cu"
#include "batch.cu"
#include "blockseq.cu"
#include "gen.cu"
#include "indices.cu"

// the sequential kernels live in their own translation units (seq_svrg.cu, seq_saga.cu, …)
int run_seq_svrg(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, double m_d);
int run_seq_saga(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, double m_d);
int run_seq_finito(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, double m_d);
int run_seq_lfinito(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, double m_d);
int run_seq_adaptive(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, double alpha, double tol_b);  // seq_adaptive.cu
int run_adaptive_init(ciao_ctx *c, const double *x0_dev, double alpha);
int run_adaptive_retry(ciao_ctx *c, int64_t i, const double *x0_dev, const double *xeps_dev, double *out_dev);
int run_adaptive_sdivg(ciao_ctx *c, const double *x0_dev, int *n_chunks_out);
int run_adaptive_av(ciao_ctx *c, const double *S_dev, const double *G_dev, double hat_gamma);
int run_block_seq(ciao_ctx *c, int alg, const int64_t *idx_prepared, int64_t K, double m_d);   // blockseq.cu
// block components (M > 1) and the complex soft-threshold take the general kernel; everything else the tuned cluster kernels
static inline bool use_block_kernel(const ciao_ctx *c) { return c->M > 1 || c->reg.kind == CIAO_REG_NORML1_PAIRS || c->force_block; }
static int run_seq(ciao_ctx *c, int alg, const int64_t *idx_prepared, int64_t K, double m_d) {
    if (use_block_kernel(c)) return run_block_seq(c, alg, idx_prepared, K, m_d);
    switch (alg) {
        case ALG_SVRG: return run_seq_svrg(c, idx_prepared, K, m_d);
        case ALG_SAGA: return run_seq_saga(c, idx_prepared, K, m_d);
        case ALG_FINITO: return run_seq_finito(c, idx_prepared, K, m_d);
        default: return run_seq_lfinito(c, idx_prepared, K, m_d);
    }
}
#include "proshi.cu"
#include "comm.cu"

void ciao_comm_destroy(ciao_ctx *c);

// ---------------------------------------------------------------------------
// small elementwise kernels
// ---------------------------------------------------------------------------
// out = prox_g(in, γ)      prox!(z, g, av, γ̂)  Finito_basic.jl:84, Finito_LFinito.jl:83
__global__ void prox_vec_kernel(const double *in, double *out, int64_t d_pad, double gamma, RegParams reg) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= d_pad) return;
    if (reg.kind == CIAO_REG_NORML1_PAIRS) {   // the thread of each element recomputes its pair (d_pad is even)
        const double2 y = prox_pair(reg, make_double2(in[j & ~1ll], in[j | 1ll]), gamma * reg.lambda, j & ~1ll);
        out[j] = (j & 1) ? y.y : y.x;
        return;
    }
    const double lo = reg.lo_v ? reg.lo_v[j] : reg.lo_s, hi = reg.hi_v ? reg.hi_v[j] : reg.hi_s;
    out[j] = prox_rt(reg.kind, in[j], gamma * reg.lambda, lo, hi);
}
// z = prox_g((1−γ)·x0, γ)   SAGA_basic.jl:48
__global__ void saga_z0_kernel(const double *x0, double *z, int64_t d_pad, double gamma, RegParams reg) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= d_pad) return;
    if (reg.kind == CIAO_REG_NORML1_PAIRS) {
        const double2 y = prox_pair(reg, make_double2(__dmul_rn(1 - gamma, x0[j & ~1ll]), __dmul_rn(1 - gamma, x0[j | 1ll])),
                                    gamma * reg.lambda, j & ~1ll);
        z[j] = (j & 1) ? y.y : y.x;
        return;
    }
    const double lo = reg.lo_v ? reg.lo_v[j] : reg.lo_s, hi = reg.hi_v ? reg.hi_v[j] : reg.hi_s;
    z[j] = prox_rt(reg.kind, __dmul_rn(1 - gamma, x0[j]), gamma * reg.lambda, lo, hi);
}
// record tails: [b_i | scale_i | γ_i | 0]
__global__ void pack_tails_kernel(double *rec, int64_t n_rows, int64_t d_pad, int64_t ld, const double *b,
                                  const double *scale, double scale_scalar) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    double *t = rec + i * ld + d_pad;
    t[0] = b[i];
    t[1] = scale ? scale[i] : scale_scalar;
    for (int k = 2; k < CIAO_TAIL; ++k) t[k] = 0.0;
}
__global__ void set_gamma_tail_kernel(double *rec, int64_t n_rows, int64_t d_pad, int64_t ld, const double *gam, double Nd,
                                      double hat_gamma, int64_t il_block, int il_rank, int il_world) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    double *t = rec + i * ld + d_pad;
    const double g = gam[il_block ? il_global(i, il_block, il_rank, il_world) : i];   // gam: first row of a contiguous shard, or all N
    t[TAIL_GAM] = g;
    t[TAIL_GAM_N] = __ddiv_rn(g, Nd);
    t[TAIL_HAT_GAM] = __ddiv_rn(hat_gamma, g);
}

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
static inline int blocks_for(int64_t n) { return (int)((n + 255) / 256); }

static void detach_peers(ciao_ctx *c);
static void free_problem(ciao_ctx *c) {
    detach_peers(c);
    cudaFree(c->rec); cudaFree(c->qd); cudaFree(c->ql); cudaFree(c->vecs); cudaFree(c->table);
    cudaFree(c->gamma_dev); cudaFree(c->partial); cudaFree(c->reg_bounds); cudaFree(c->ss); cudaFree(c->adapt); cudaFree(c->adapt_scal); cudaFree(c->adapt_counters); cudaFree(c->gpair);
    c->rec = c->qd = c->ql = c->vecs = c->table = c->gamma_dev = c->partial = c->reg_bounds = c->ss = c->adapt = c->adapt_scal = c->gpair = nullptr;
    c->ss_cap = 0;
    c->adapt_counters = nullptr;
    c->reg = RegParams{CIAO_REG_ZERO, 0, 0, 0, nullptr, nullptr};
    c->loss_kind = -1;
    c->algo = 0;
}

static int alloc_common(ciao_ctx *c, int64_t N_total, int64_t row0, int64_t n_rows, int64_t d) {
    free_problem(c);
    c->N_total = N_total; c->row0 = row0; c->n_rows = n_rows; c->d = d;
    c->M = 1;
    c->win0 = c->win_n = 0;
    c->cz_valid = false; c->cz_local_valid = false;
    c->d_pad = (d + 3) / 4 * 4;
    c->ld = c->d_pad + CIAO_TAIL;
    CUDA_TRY(cudaMalloc(&c->vecs, (size_t)CIAO_NUM_VECS * c->d_pad * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(c->vecs, 0, (size_t)CIAO_NUM_VECS * c->d_pad * sizeof(double), c->stream));
    CUDA_TRY(cudaMalloc(&c->partial, (size_t)(c->d_pad + 8) * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(c->partial, 0, (size_t)(c->d_pad + 8) * sizeof(double), c->stream));
    return CIAO_OK;
}

static bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Pinned (page-locked / registered) host memory: cudaMemcpyAsync from it is truly asynchronous, so a borrowed buffer must be
// synchronised before the call returns.  Pageable memory is staged by the driver before cudaMemcpyAsync returns (CUDA runtime
// "API synchronization behavior"), so no extra stream synchronisation is needed — which keeps small calls (the one-step-per-
// call iterator protocol) from serialising on the previous step's kernel.
static bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;   // unknown: be conservative
    }
    return at.type != cudaMemoryTypeUnregistered;
}

// host vector (length d) → state vector `which` (padding stays zero); blocks until the copy is done
static int upload_vec(ciao_ctx *c, int which, const double *x) {
    double *dst = ctx_vec(c, which);
    if (is_device_ptr(x)) {
        CUDA_TRY(cudaMemcpyAsync(dst, x, (size_t)c->d * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    } else {
        CUDA_TRY(cudaMemcpyAsync(dst, x, (size_t)c->d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return CIAO_OK;
}

static int copy_vec(ciao_ctx *c, int dst, int src) {
    CUDA_TRY(cudaMemcpyAsync(ctx_vec(c, dst), ctx_vec(c, src), (size_t)c->d_pad * sizeof(double), cudaMemcpyDeviceToDevice,
                             c->stream));
    return CIAO_OK;
}

static int check_err_flag(ciao_ctx *c) {
    int h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, c->err_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h) {
        CUDA_TRY(cudaMemsetAsync(c->err_dev, 0, sizeof(int), c->stream));
        if (h == 3) CIAO_FAIL(CIAO_ERR_COMM, "peer exchange timed out: a rank of the collective did not arrive (ciao_comm_p2p_attach)");
        CIAO_FAIL(CIAO_ERR_INVALID, "index out of range 1..N (or malformed batch_ptr / batch_order) in a previous call; the steps of that "
                  "call were not executed");
    }
    return CIAO_OK;
}

static int reserve_idx(ciao_ctx *c, size_t n) {
    if (n > c->idx_cap) {
        // keep staged raw indices across a growth of the buffers
        int64_t *nr = nullptr, *np = nullptr;
        const size_t cap = std::max(n, c->idx_cap * 2);
        CUDA_TRY(cudaMalloc(&nr, cap * sizeof(int64_t)));
        if (cudaMalloc(&np, cap * sizeof(int64_t)) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(nr);
            CIAO_FAIL(CIAO_ERR_OOM, "index staging buffers: out of device memory (%zu indices)", cap);
        }
        if (c->idx_raw && c->staged > 0)
            CUDA_TRY(cudaMemcpyAsync(nr, c->idx_raw, (size_t)c->staged * sizeof(int64_t), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->idx_raw); cudaFree(c->idx_prep);
        c->idx_raw = nr; c->idx_prep = np; c->idx_cap = cap;
    }
    return CIAO_OK;
}

// Brings n raw (1-based) indices to the device: *raw_dev points at them.
static int fetch_raw_indices(ciao_ctx *c, const int64_t *idx, int64_t n, const int64_t **raw_dev) {
    if (n < 0) CIAO_FAIL(CIAO_ERR_INVALID, "negative index count");
    if (idx == nullptr) {
        if (c->staged < n) CIAO_FAIL(CIAO_ERR_STATE, "idx == NULL but only %lld indices are staged (need %lld)",
                                     (long long)c->staged, (long long)n);
        *raw_dev = c->idx_raw;
        return CIAO_OK;
    }
    CIAO_TRY(reserve_idx(c, (size_t)std::max<int64_t>(n, 1)));
    if (is_device_ptr(idx)) {
        *raw_dev = idx;
        return CIAO_OK;
    }
    c->staged = 0;
    const bool pinned = host_ptr_is_pinned(idx);
    CUDA_TRY(cudaMemcpyAsync(c->idx_raw, idx, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    if (pinned) CUDA_TRY(cudaStreamSynchronize(c->stream));  // host buffer is borrowed for the call only
    *raw_dev = c->idx_raw;
    return CIAO_OK;
}

static int upload_ptr(ciao_ctx *c, const int64_t *ptr, int64_t n, const int64_t **out) {
    if (is_device_ptr(ptr)) {
        *out = ptr;
        return CIAO_OK;
    }
    if ((size_t)n > c->ptr_cap) {
        cudaFree(c->ptr_dev);
        c->ptr_dev = nullptr;
        c->ptr_cap = 0;
        CUDA_TRY(cudaMalloc(&c->ptr_dev, (size_t)n * 2 * sizeof(int64_t)));
        c->ptr_cap = (size_t)n * 2;
    }
    CUDA_TRY(cudaMemcpyAsync(c->ptr_dev, ptr, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out = c->ptr_dev;
    return CIAO_OK;
}

static int need_rows(ciao_ctx *c, const char *who, bool whole) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "%s: null context", who);
    if (c->loss_kind != CIAO_LOSS_LS && c->loss_kind != CIAO_LOSS_LOGISTIC)
        CIAO_FAIL(CIAO_ERR_STATE, "%s: no row problem set (ciao_set_rows / ciao_gen_synthetic first)", who);
    if (whole && c->n_rows != c->N_total && c->peers.n <= 1)
        CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "%s: the sequential loops need all N rows: this context holds a shard and no peer rows are attached "
                  "(ciao_attach_peer_rows)", who);
    CUDA_TRY(cudaSetDevice(c->device));
    return CIAO_OK;
}

// index staging buffers sized once per solver (cudaMalloc/cudaFree next to a 137 GB allocation cost ~100 ms each)
static int reserve_idx(ciao_ctx *c, size_t n);
static int reserve_for_solver(ciao_ctx *c) { return reserve_idx(c, (size_t)std::max<int64_t>(c->N_total, 1 << 16)); }

// The N×d table lives next to the rows it belongs to: a row shard holds the table rows of its shard (the table-init passes
// shard like the full gradient); the sequential steps need the whole table on one GPU (need_whole_table).
static int alloc_table(ciao_ctx *c) {
    if (!c->table) CUDA_TRY(cudaMalloc(&c->table, (size_t)c->n_rows * c->d_pad * sizeof(double)));
    return CIAO_OK;
}
static int need_whole_table(ciao_ctx *c, const char *who) {
    if (c->n_rows != c->N_total)
        CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "%s: the N×d tables of SAGA/Finito are not sharded for the sequential steps — this context holds "
                  "rows %lld..%lld of %lld (table init passes do shard)", who, (long long)c->row0, (long long)(c->row0 + c->n_rows),
                  (long long)c->N_total);
    return CIAO_OK;
}

static int set_gammas(ciao_ctx *c, const double *gamma_N, bool tails) {
    if (!c->gamma_dev) CUDA_TRY(cudaMalloc(&c->gamma_dev, (size_t)c->N_total * 2 * sizeof(double)));  // γ_i, then γ_i/N
    if (is_device_ptr(gamma_N)) {
        CUDA_TRY(cudaMemcpyAsync(c->gamma_dev, gamma_N, (size_t)c->N_total * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    } else {
        CUDA_TRY(cudaMemcpyAsync(c->gamma_dev, gamma_N, (size_t)c->N_total * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    if (tails && c->M == 1) {   // block components read γ_i from gamma_dev (blockseq.cu)
        set_gamma_tail_kernel<<<blocks_for(c->n_rows), 256, 0, c->stream>>>(c->rec, c->n_rows, c->d_pad, c->ld,
                                                                             c->gamma_dev + (c->il_block ? 0 : c->row0), (double)c->N_total,
                                                                             c->hat_gamma, c->il_block, c->il_rank, c->il_world);  // γ of MY rows
        CUDA_TRY(cudaGetLastError());
        c->timing.launches += 1;
    }
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------
extern "C" int ciao_version(void) { return CIAO_VERSION; }
extern "C" const char *ciao_last_error(void) { return g_err; }

extern "C" int ciao_device_count(int *n) {
    if (!n) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_device_count: null output");
    *n = 0;
    CUDA_TRY(cudaGetDeviceCount(n));
    return CIAO_OK;
}

extern "C" int ciao_create(ciao_ctx **out, int device) {
    if (!out) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_create: null output");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        CIAO_FAIL(CIAO_ERR_CUDA, "no CUDA device available (%s); libciao_cuda has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_create: device %d not in [0,%d)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    ciao_ctx *c = new (std::nothrow) ciao_ctx();
    if (!c) CIAO_FAIL(CIAO_ERR_OOM, "host allocation failed");
    c->device = device;
    c->cache_cz = !(getenv("CIAO_CACHE_CZ") != nullptr && getenv("CIAO_CACHE_CZ")[0] == '0');
    c->seq_table_ldg = getenv("CIAO_SEQ_TABLE_LDG") != nullptr && getenv("CIAO_SEQ_TABLE_LDG")[0] == '1';
    c->force_block = getenv("CIAO_FORCE_BLOCK_KERNEL") != nullptr && getenv("CIAO_FORCE_BLOCK_KERNEL")[0] == '1';
    c->batch_persistent = !(getenv("CIAO_BATCH_PER_LAUNCH") != nullptr && getenv("CIAO_BATCH_PER_LAUNCH")[0] == '1');
    c->num_sms = prop.multiProcessorCount;
    if (const char *pos = getenv("CIAO_SEQ_CLUSTER_POS")) c->seq_cluster_pos = atoi(pos);
    auto init_device_objects = [&]() -> int {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        cudaEvent_t *evs[] = {&c->ev_pa, &c->ev_pb, &c->ev_sa, &c->ev_sb, &c->tm_a, &c->tm_b, &c->ev_pc};
        for (auto ev : evs) CUDA_TRY(cudaEventCreate(ev));
        CUDA_TRY(cudaMalloc(&c->err_dev, sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(c->err_dev, 0, sizeof(int), c->stream));
        CUDA_TRY(cudaMalloc(&c->seq_smid, 16 * sizeof(int) + 2 * sizeof(long long)));   // SM ids, then {cycles, ns} of the last kernel
        CUDA_TRY(cudaMemsetAsync(c->seq_smid, 0xff, 16 * sizeof(int), c->stream));
        CUDA_TRY(cudaMemsetAsync(c->seq_smid + 16, 0, 2 * sizeof(long long), c->stream));
        return CIAO_OK;
    };
    const int rc = init_device_objects();
    if (rc != CIAO_OK) {  // a half-built context is torn down here (the error text is already set), never handed out
        ciao_destroy(c);
        return rc;
    }
    *out = c;
    return CIAO_OK;
}

extern "C" int ciao_destroy(ciao_ctx *c) {
    if (!c) return CIAO_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    ciao_comm_destroy(c);
    free_problem(c);
    cudaFree(c->ws); cudaFree(c->idx_raw); cudaFree(c->idx_prep); cudaFree(c->ptr_dev); cudaFree(c->err_dev); cudaFree(c->grid_bar); cudaFree(c->ll_buf); cudaFree(c->seq_smid);
    cudaEvent_t evs[] = {c->ev_pa, c->ev_pb, c->ev_sa, c->ev_sb, c->tm_a, c->tm_b, c->ev_pc};
    for (auto ev : evs) if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return CIAO_OK;
}

extern "C" int ciao_sync(ciao_ctx *c) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_sync: null context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_err_flag(c);
}

extern "C" int ciao_set_tuning(ciao_ctx *c, int pass_threads, int pass_stages, int pass_ctas_per_sm, int seq_cluster, int seq_threads) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_tuning: null context");
    if (seq_cluster != 0 && seq_cluster != 1 && seq_cluster != 2 && seq_cluster != 4 && seq_cluster != 8 && seq_cluster != 16)
        CIAO_FAIL(CIAO_ERR_INVALID, "seq_cluster must be 0, 1, 2, 4, 8 or 16");
    if (pass_threads < 0 || pass_threads > 512 || seq_threads < 0 || seq_threads > 512 || pass_stages < 0 || pass_ctas_per_sm < 0 || pass_ctas_per_sm > 8)
        CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_tuning: value out of range");
    c->pass_threads = pass_threads; c->pass_stages = pass_stages; c->pass_ctas = pass_ctas_per_sm;
    c->seq_cluster = seq_cluster; c->seq_threads = seq_threads;
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
// remote rows over NVLink: every rank maps the other ranks' row shards with CUDA IPC, so
// the sequential inner epochs (SVRG, LFinito) can sample from a problem that is sharded over several GPUs.
// ---------------------------------------------------------------------------
static void detach_peers(ciao_ctx *c) {
    for (int s = 0; s < CIAO_MAX_PEERS; ++s) {
        if (c->peer_mapped[s]) cudaIpcCloseMemHandle(c->peer_mapped[s]);
        c->peer_mapped[s] = nullptr;
    }
    c->peers.n = 0;
}

extern "C" int ciao_rows_ipc_handle(ciao_ctx *c, void *out64) {
    if (!c || !out64) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_rows_ipc_handle: null argument");
    if (!c->rec) CIAO_FAIL(CIAO_ERR_STATE, "ciao_rows_ipc_handle: no row problem set");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->rec));
    memcpy(out64, &h, sizeof(h));
    return CIAO_OK;
}

extern "C" int ciao_attach_peer_rows(ciao_ctx *c, int n_shards, const void *handles64, const int64_t *row0, const int64_t *n_rows,
                                     int my_shard) {
    if (!c || !handles64 || !row0 || !n_rows) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_attach_peer_rows: null argument");
    if (c->il_block) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "ciao_attach_peer_rows: remote rows are for contiguous shards (interleaved shards serve minibatches and passes)");
    if (n_shards < 1 || n_shards > CIAO_MAX_PEERS || my_shard < 0 || my_shard >= n_shards)
        CIAO_FAIL(CIAO_ERR_INVALID, "ciao_attach_peer_rows: 1..%d shards, my_shard inside", CIAO_MAX_PEERS);
    if (!c->rec) CIAO_FAIL(CIAO_ERR_STATE, "ciao_attach_peer_rows: no row problem set");
    if (row0[my_shard] != c->row0 || n_rows[my_shard] != c->n_rows) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_attach_peer_rows: my shard does not match this context's rows");
    int64_t next = 0;
    for (int s = 0; s < n_shards; ++s) {
        if (row0[s] != next || n_rows[s] <= 0) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_attach_peer_rows: shards must tile [0, N) contiguously in order");
        next += n_rows[s];
    }
    if (next != c->N_total) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_attach_peer_rows: shards cover %lld rows, N = %lld", (long long)next, (long long)c->N_total);
    CUDA_TRY(cudaSetDevice(c->device));
    detach_peers(c);
    for (int s = 0; s < n_shards; ++s) {
        c->peers.start[s] = row0[s];
        if (s == my_shard) {
            c->peers.base[s] = c->rec;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles64 + 64 * (size_t)s, sizeof(h));
        void *ptr = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_mapped[s] = ptr;
        c->peers.base[s] = (const double *)ptr;
    }
    c->peers.start[n_shards] = c->N_total;
    c->peers.n = n_shards;
    return CIAO_OK;
}

extern "C" int ciao_set_pass_window(ciao_ctx *c, int64_t row_lo, int64_t n) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_pass_window: null context");
    c->win_uniform = false;
    c->cz_valid = false; c->cz_local_valid = false;
    if (n == 0) {
        c->win0 = c->win_n = 0;
        if (c->world <= 1) return CIAO_OK;
    } else {
        if (row_lo < 0 || n < 0 || row_lo + n > c->n_rows) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_pass_window: window outside the local rows");
        c->win0 = row_lo;
        c->win_n = n;
    }
    if (c->world > 1 && c->nccl_comm && c->partial) {
        // collective when a communicator exists: do all ranks window the same replicated problem uniformly?  Then a pass can
        // all-gather the per-row step scalars of the windows for the replicated inner epochs.
        CUDA_TRY(cudaSetDevice(c->device));
        const bool ok = c->win_n > 0 && c->n_rows == c->N_total && c->N_total % c->world == 0 && c->win_n == c->N_total / c->world &&
                        c->win0 == (int64_t)c->rank * c->win_n;
        double flag = ok ? 0.0 : 1.0;
        double *slot = c->partial + c->d_pad + 4;
        CUDA_TRY(cudaMemcpyAsync(slot, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CIAO_TRY(ciao_comm_allreduce(c, slot, 1, 1));
        CUDA_TRY(cudaMemcpyAsync(&flag, slot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->win_uniform = flag == 0.0;
    }
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
// problem
// ---------------------------------------------------------------------------
// A (n_rows × d, leading dimension lda) → the a_i part of the row records.  Device and pinned sources go in one 2-D copy.
// A large PAGEABLE source (a Julia Matrix, a numpy array: the case of `solver(x0; F = …)`) would be staged by the driver
// through its own bounce buffer at ≈ 10 GB/s (measured, 34 GB); instead host threads copy row chunks into two pinned buffers
// while the previous chunk is in flight over PCIe.
static int upload_rows(ciao_ctx *c, const double *A, int64_t lda, int64_t n_rows, int64_t d) {
    const size_t row_bytes = (size_t)d * sizeof(double);
    const size_t total = row_bytes * (size_t)n_rows;
    const bool pageable = !is_device_ptr(A) && !host_ptr_is_pinned(A);
    if (!pageable || total < ((size_t)256 << 20)) {
        CUDA_TRY(cudaMemcpy2DAsync(c->rec, (size_t)c->ld * sizeof(double), A, (size_t)lda * sizeof(double), row_bytes, (size_t)n_rows,
                                   cudaMemcpyDefault, c->stream));
        return CIAO_OK;
    }
    const size_t chunk_bytes = (size_t)128 << 20;
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(chunk_bytes / row_bytes));
    double *pin[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    int rc = CIAO_OK;
    auto cleanup = [&]() {
        for (int b = 0; b < 2; ++b) {
            if (pin[b]) cudaFreeHost(pin[b]);
            if (done[b]) cudaEventDestroy(done[b]);
        }
    };
    for (int b = 0; b < 2; ++b) {
        if (cudaHostAlloc(&pin[b], (size_t)chunk_rows * row_bytes, cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            cleanup();   // no pinned memory to be had: the plain copy still works
            CUDA_TRY(cudaMemcpy2DAsync(c->rec, (size_t)c->ld * sizeof(double), A, (size_t)lda * sizeof(double), row_bytes, (size_t)n_rows,
                                       cudaMemcpyDefault, c->stream));
            return CIAO_OK;
        }
    }
    const int n_thr = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
    int64_t r0 = 0;
    for (int it = 0; r0 < n_rows; ++it, r0 += chunk_rows) {
        const int b = it & 1;
        const int64_t nr = std::min(chunk_rows, n_rows - r0);
        if (it >= 2 && cudaEventSynchronize(done[b]) != cudaSuccess) { rc = CIAO_ERR_CUDA; break; }   // the buffer's last DMA is over
        std::vector<std::thread> th;
        for (int t = 0; t < n_thr; ++t)
            th.emplace_back([&, t]() {
                for (int64_t r = nr * t / n_thr; r < nr * (t + 1) / n_thr; ++r)
                    memcpy(pin[b] + (size_t)r * d, A + (size_t)(r0 + r) * lda, row_bytes);
            });
        for (auto &x : th) x.join();
        if (cudaMemcpy2DAsync(c->rec + (size_t)r0 * c->ld, (size_t)c->ld * sizeof(double), pin[b], row_bytes, row_bytes, (size_t)nr,
                              cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaEventRecord(done[b], c->stream) != cudaSuccess) { rc = CIAO_ERR_CUDA; break; }
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = CIAO_ERR_CUDA;
    if (rc != CIAO_OK) ciao_set_error("ciao_set_rows: staged upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    cleanup();
    return rc;
}

static int check_interleave(ciao_ctx *c, const char *who, int64_t N_total, int64_t row0, int64_t n_rows);
extern "C" int ciao_set_rows(ciao_ctx *c, int loss_kind, int64_t N_total, int64_t row0, int64_t n_rows, int64_t d,
                             const double *A, int64_t lda, const double *b_or_y, const double *scale, double scale_scalar) {
    if (!c || !A || !b_or_y) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_rows: null argument");
    if (loss_kind != CIAO_LOSS_LS && loss_kind != CIAO_LOSS_LOGISTIC) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "ciao_set_rows: unknown loss kind %d", loss_kind);
    if (N_total <= 0 || n_rows <= 0 || d <= 0 || lda < d || row0 < 0 || row0 + n_rows > N_total)
        CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_rows: bad shape N=%lld row0=%lld n_rows=%lld d=%lld lda=%lld", (long long)N_total,
                  (long long)row0, (long long)n_rows, (long long)d, (long long)lda);
    CUDA_TRY(cudaSetDevice(c->device));
    CIAO_TRY(check_interleave(c, "ciao_set_rows", N_total, row0, n_rows));
    CIAO_TRY(alloc_common(c, N_total, row0, n_rows, d));
    CUDA_TRY(cudaMalloc(&c->rec, (size_t)n_rows * c->ld * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(c->rec, 0, (size_t)n_rows * c->ld * sizeof(double), c->stream));
    CIAO_TRY(upload_rows(c, A, lda, n_rows, d));
    double *tmp = nullptr;
    CUDA_TRY(cudaMalloc(&tmp, (size_t)n_rows * 2 * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(tmp, b_or_y, (size_t)n_rows * sizeof(double), cudaMemcpyDefault, c->stream));
    if (scale) CUDA_TRY(cudaMemcpyAsync(tmp + n_rows, scale, (size_t)n_rows * sizeof(double), cudaMemcpyDefault, c->stream));
    pack_tails_kernel<<<blocks_for(n_rows), 256, 0, c->stream>>>(c->rec, n_rows, c->d_pad, c->ld, tmp, scale ? tmp + n_rows : nullptr, scale_scalar);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(tmp);
    c->loss_kind = loss_kind;
    return CIAO_OK;
}

// per-row tails of a block problem: row r of component i = r / M carries b_r and the component's λ_i
__global__ void pack_block_tails_kernel(double *rec, int64_t n_rows_total, int64_t M, int64_t d_pad, int64_t ld, const double *b,
                                        const double *scale, double scale_scalar) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows_total) return;
    double *t = rec + r * ld + d_pad;
    t[0] = b[r];
    t[1] = scale ? scale[r / M] : scale_scalar;
    for (int k = 2; k < CIAO_TAIL; ++k) t[k] = 0.0;
}

extern "C" int ciao_set_row_blocks(ciao_ctx *c, int loss_kind, int64_t N, int64_t M, int64_t d, const double *A, int64_t lda,
                                   const double *b_or_y, const double *scale, double scale_scalar) {
    if (!c || !A || !b_or_y) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_row_blocks: null argument");
    if (loss_kind != CIAO_LOSS_LS && loss_kind != CIAO_LOSS_LOGISTIC) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "ciao_set_row_blocks: unknown loss kind %d", loss_kind);
    if (N <= 0 || M <= 0 || M > 4096 || d <= 0 || lda < d) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_row_blocks: bad shape N=%lld M=%lld d=%lld lda=%lld",
                                                                     (long long)N, (long long)M, (long long)d, (long long)lda);
    if (d > 8192) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "ciao_set_row_blocks: d = %lld exceeds 8192", (long long)d);
    CUDA_TRY(cudaSetDevice(c->device));
    CIAO_TRY(alloc_common(c, N, 0, N, d));
    c->M = (int)M;
    const int64_t rows = N * M;
    CUDA_TRY(cudaMalloc(&c->rec, (size_t)rows * c->ld * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(c->rec, 0, (size_t)rows * c->ld * sizeof(double), c->stream));
    CUDA_TRY(cudaMemcpy2DAsync(c->rec, (size_t)c->ld * sizeof(double), A, (size_t)lda * sizeof(double), (size_t)d * sizeof(double),
                               (size_t)rows, cudaMemcpyDefault, c->stream));
    double *tmp = nullptr;
    CUDA_TRY(cudaMalloc(&tmp, (size_t)(rows + N) * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(tmp, b_or_y, (size_t)rows * sizeof(double), cudaMemcpyDefault, c->stream));
    if (scale) CUDA_TRY(cudaMemcpyAsync(tmp + rows, scale, (size_t)N * sizeof(double), cudaMemcpyDefault, c->stream));
    pack_block_tails_kernel<<<blocks_for(rows), 256, 0, c->stream>>>(c->rec, rows, M, c->d_pad, c->ld, tmp, scale ? tmp + rows : nullptr, scale_scalar);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(tmp);
    c->loss_kind = loss_kind;
    return CIAO_OK;
}

extern "C" int ciao_set_blocks(ciao_ctx *c, int64_t N, int64_t n, const double *Qdiag, int64_t ldq, const double *qlin,
                               int64_t ldl, double box_lo, double box_hi, double eta) {
    if (!c || !Qdiag || !qlin) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_blocks: null argument");
    if (N <= 0 || n <= 0 || ldq < n || ldl < n || !(box_lo <= box_hi)) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_blocks: bad shape");
    CUDA_TRY(cudaSetDevice(c->device));
    CIAO_TRY(alloc_common(c, N, 0, N, n));
    const size_t bytes = (size_t)N * c->d_pad * sizeof(double);
    CUDA_TRY(cudaMalloc(&c->qd, bytes));
    CUDA_TRY(cudaMalloc(&c->ql, bytes));
    CUDA_TRY(cudaMemsetAsync(c->qd, 0, bytes, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->ql, 0, bytes, c->stream));
    CUDA_TRY(cudaMemcpy2DAsync(c->qd, (size_t)c->d_pad * 8, Qdiag, (size_t)ldq * 8, (size_t)n * 8, (size_t)N, cudaMemcpyDefault, c->stream));
    CUDA_TRY(cudaMemcpy2DAsync(c->ql, (size_t)c->d_pad * 8, qlin, (size_t)ldl * 8, (size_t)n * 8, (size_t)N, cudaMemcpyDefault, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->box_lo = box_lo; c->box_hi = box_hi; c->eta = eta;
    c->loss_kind = CIAO_LOSS_DIAGQUAD;
    return CIAO_OK;
}

extern "C" int ciao_set_reg(ciao_ctx *c, int reg_kind, const double *params, int64_t nparams) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_reg: null context");
    if (c->loss_kind < 0) CIAO_FAIL(CIAO_ERR_STATE, "ciao_set_reg: set the problem (F) first");
    CUDA_TRY(cudaSetDevice(c->device));
    RegParams r{reg_kind, 0, 0, 0, nullptr, nullptr};
    if (reg_kind == CIAO_REG_ZERO) {
    } else if (reg_kind == CIAO_REG_NORML1 || reg_kind == CIAO_REG_NORML1_PAIRS) {
        if (!params || nparams != 1 || !(params[0] >= 0)) CIAO_FAIL(CIAO_ERR_INVALID, "NormL1 takes one parameter λ ≥ 0");
        if (reg_kind == CIAO_REG_NORML1_PAIRS && (c->d % 2) != 0) CIAO_FAIL(CIAO_ERR_INVALID, "NormL1 on (re, im) pairs needs an even d");
        r.lambda = params[0];
    } else if (reg_kind == CIAO_REG_INDBOX) {
        if (params && nparams == 2) {
            r.lo_s = params[0]; r.hi_s = params[1];
            if (!(r.lo_s <= r.hi_s)) CIAO_FAIL(CIAO_ERR_INVALID, "IndBox: lo > hi");
        } else if (params && nparams == 2 * c->d) {
            cudaFree(c->reg_bounds);
            c->reg_bounds = nullptr;
            CUDA_TRY(cudaMalloc(&c->reg_bounds, (size_t)2 * c->d_pad * sizeof(double)));
            std::vector<double> h((size_t)2 * c->d_pad, 0.0);
            for (int64_t j = 0; j < c->d; ++j) {
                if (!(params[j] <= params[c->d + j])) CIAO_FAIL(CIAO_ERR_INVALID, "IndBox: lo[%lld] > hi[%lld]", (long long)j, (long long)j);
                h[j] = params[j];
                h[c->d_pad + j] = params[c->d + j];
            }
            CUDA_TRY(cudaMemcpy(c->reg_bounds, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
            r.lo_v = c->reg_bounds;
            r.hi_v = c->reg_bounds + c->d_pad;
        } else {
            CIAO_FAIL(CIAO_ERR_INVALID, "IndBox takes {lo,hi} or lo[d] followed by hi[d]");
        }
    } else {
        CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "unknown g kind %d (supported: Zero, NormL1, IndBox)", reg_kind);
    }
    c->reg = r;
    return CIAO_OK;
}

// Interleaved row shards: declared before the rows arrive.  The data calls then take row0 = rank·block_rows (the first row this
// rank owns) and n_rows = the number of rows it owns.
extern "C" int ciao_set_row_interleave(ciao_ctx *c, int64_t block_rows, int rank, int world) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_row_interleave: null context");
    if (c->rec || c->qd) CIAO_FAIL(CIAO_ERR_STATE, "ciao_set_row_interleave: call it before the rows are set");
    if (block_rows == 0) { c->il_block = 0; c->il_rank = 0; c->il_world = 1; return CIAO_OK; }
    if (block_rows < 1 || world < 1 || world > CIAO_MAX_PEERS || rank < 0 || rank >= world)
        CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_row_interleave: block_rows ≥ 1, 1..%d ranks, rank inside", CIAO_MAX_PEERS);
    c->il_block = block_rows; c->il_rank = rank; c->il_world = world;
    return CIAO_OK;
}
static int check_interleave(ciao_ctx *c, const char *who, int64_t N_total, int64_t row0, int64_t n_rows) {
    if (!c->il_block) return CIAO_OK;
    const int64_t want = il_count(N_total, c->il_block, c->il_rank, c->il_world);
    if (row0 != (int64_t)c->il_rank * c->il_block || n_rows != want)
        CIAO_FAIL(CIAO_ERR_INVALID, "%s: interleaved shard of rank %d/%d with blocks of %lld rows holds %lld rows starting at row %lld",
                  who, c->il_rank, c->il_world, (long long)c->il_block, (long long)want, (long long)((int64_t)c->il_rank * c->il_block));
    return CIAO_OK;
}
extern "C" int ciao_gen_synthetic(ciao_ctx *c, int kind, int64_t N_total, int64_t row0, int64_t n_rows, int64_t d,
                                  uint64_t seed, double scale) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_gen_synthetic: null context");
    if (kind < 0 || kind > 2 || N_total <= 0 || n_rows <= 0 || d <= 0 || row0 < 0 || row0 + n_rows > N_total)
        CIAO_FAIL(CIAO_ERR_INVALID, "ciao_gen_synthetic: bad arguments");
    CIAO_TRY(check_interleave(c, "ciao_gen_synthetic", N_total, row0, n_rows));
    CUDA_TRY(cudaSetDevice(c->device));
    CIAO_TRY(alloc_common(c, N_total, row0, n_rows, d));
    if (kind == CIAO_SYNTH_SHARING) {
        if (row0 != 0 || n_rows != N_total) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "sharing blocks are not sharded");
        const size_t bytes = (size_t)N_total * c->d_pad * sizeof(double);
        CUDA_TRY(cudaMalloc(&c->qd, bytes));
        CUDA_TRY(cudaMalloc(&c->ql, bytes));
        CIAO_TRY(launch_gen_blocks(c, seed));
        c->box_lo = -2.0; c->box_hi = 2.0; c->eta = 10.0 * (double)N_total;  // test_sharing.jl:15-16
        c->loss_kind = CIAO_LOSS_DIAGQUAD;
    } else {
        CUDA_TRY(cudaMalloc(&c->rec, (size_t)n_rows * c->ld * sizeof(double)));
        CIAO_TRY(launch_gen_records(c, kind, seed, scale));
        c->loss_kind = kind == CIAO_SYNTH_LASSO ? CIAO_LOSS_LS : CIAO_LOSS_LOGISTIC;
    }
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
// passes
// ---------------------------------------------------------------------------
// one streaming pass with its closing update out = base + scale·(Σ/den) fused into the pass's tail kernel (pass.cu)
static int run_pass_finish(ciao_ctx *c, int mode, const double *x_dev, bool cache_cz, const double *base, double scale, double den,
                           double *out) {
    const FinishSpec f{base, scale, den, out};
    return run_row_pass(c, mode, x_dev, cache_cz, &f);
}

extern "C" int ciao_full_gradient(ciao_ctx *c, const double *x, double scale, double *out) {
    CIAO_TRY(need_rows(c, "ciao_full_gradient", false));
    if (x) CIAO_TRY(upload_vec(c, CIAO_VEC_X, x));
    CIAO_TRY(run_pass_finish(c, PASS_GRAD, ctx_vec(c, CIAO_VEC_X), false, nullptr, scale, 1.0, ctx_vec(c, CIAO_VEC_TMP)));
    if (out) {
        CUDA_TRY(cudaMemcpyAsync(out, ctx_vec(c, CIAO_VEC_TMP), (size_t)c->d * sizeof(double), cudaMemcpyDefault, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return CIAO_OK;
}

extern "C" int ciao_objective(ciao_ctx *c, const double *x, double *f_mean, double *g_val) {
    CIAO_TRY(need_rows(c, "ciao_objective", false));
    if (!x || !f_mean) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_objective: null argument");
    std::vector<double> hx((size_t)c->d);
    CUDA_TRY(cudaMemcpy(hx.data(), x, (size_t)c->d * sizeof(double), cudaMemcpyDefault));
    CIAO_TRY(upload_vec(c, CIAO_VEC_X, hx.data()));
    CIAO_TRY(run_row_pass(c, PASS_GRAD, ctx_vec(c, CIAO_VEC_X)));
    double fs = 0.0;
    CUDA_TRY(cudaMemcpyAsync(&fs, c->partial + c->d_pad, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *f_mean = fs / (double)c->N_total;
    if (g_val) {
        double g = 0.0;
        if (c->reg.kind == CIAO_REG_NORML1) {
            for (int64_t j = 0; j < c->d; ++j) g += fabs(hx[j]);
            g *= c->reg.lambda;
        } else if (c->reg.kind == CIAO_REG_NORML1_PAIRS) {   // λ Σ_k |x_k| over the complex entries
            for (int64_t j = 0; j + 1 < c->d; j += 2) g += hypot(hx[j], hx[j + 1]);
            g *= c->reg.lambda;
        } else if (c->reg.kind == CIAO_REG_INDBOX) {   // IndBox: 0 inside [lo, hi], +Inf outside (ProximalOperators indBox.jl call operator)
            std::vector<double> bounds;
            if (c->reg.lo_v) {
                bounds.resize((size_t)2 * c->d_pad);
                CUDA_TRY(cudaMemcpy(bounds.data(), c->reg_bounds, bounds.size() * sizeof(double), cudaMemcpyDeviceToHost));
            }
            for (int64_t j = 0; j < c->d; ++j) {
                const double lo = c->reg.lo_v ? bounds[j] : c->reg.lo_s, hi = c->reg.lo_v ? bounds[c->d_pad + j] : c->reg.hi_s;
                if (hx[j] < lo || hx[j] > hi) {
                    g = INFINITY;
                    break;
                }
            }
        }
        *g_val = g;
    }
    return CIAO_OK;
}

extern "C" int ciao_max_row_sqnorm(ciao_ctx *c, double *out) {
    CIAO_TRY(need_rows(c, "ciao_max_row_sqnorm", false));
    if (!out) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_max_row_sqnorm: null output");
    CIAO_TRY(run_row_pass(c, PASS_NORMS, ctx_vec(c, CIAO_VEC_X)));
    CUDA_TRY(cudaMemcpyAsync(out, c->partial + c->d_pad, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
// SVRG / SVRG++
// ---------------------------------------------------------------------------
extern "C" int ciao_svrg_init(ciao_ctx *c, const double *x0, double gamma, int plus) {
    CIAO_TRY(need_rows(c, "ciao_svrg_init", false));
    if (!x0 || !(gamma > 0)) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_svrg_init: x0 is null or γ ≤ 0 (SVRG.jl:39)");
    c->algo = ALG_SVRG; c->gamma = gamma; c->plus = plus ? 1 : 0;
    CIAO_TRY(reserve_for_solver(c));
    CIAO_TRY(upload_vec(c, CIAO_VEC_Z_FULL, x0));                                  // z_full = copy(x0)  :64
    CIAO_TRY(copy_vec(c, CIAO_VEC_W, CIAO_VEC_Z_FULL));                            // w = copy(x0)       :66
    CUDA_TRY(cudaMemsetAsync(ctx_vec(c, CIAO_VEC_Z), 0, (size_t)c->d_pad * 8, c->stream));  // z = 0        :65
    return run_pass_finish(c, PASS_GRAD, ctx_vec(c, CIAO_VEC_Z_FULL), c->cache_cz, nullptr, 1.0, (double)c->N_total,
                           ctx_vec(c, CIAO_VEC_AV));                                  // :58-63
}

extern "C" int ciao_svrg_epoch(ciao_ctx *c, const int64_t *idx, int64_t m) {
    CIAO_TRY(need_rows(c, "ciao_svrg_epoch", true));
    if (c->algo != ALG_SVRG) CIAO_FAIL(CIAO_ERR_STATE, "ciao_svrg_epoch before ciao_svrg_init");
    if (m <= 0) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_svrg_epoch: m must be positive");
    const int64_t *raw;
    CIAO_TRY(fetch_raw_indices(c, idx, m, &raw));
    CIAO_TRY(launch_prep_indices(c, raw, m, c->N_total, c->idx_prep));
    CIAO_TRY(run_seq(c, ALG_SVRG, c->idx_prep, m, (double)m));                     // :73-86
    return run_pass_finish(c, PASS_GRAD, ctx_vec(c, CIAO_VEC_Z_FULL), c->cache_cz, nullptr, 1.0, (double)c->N_total,
                           ctx_vec(c, CIAO_VEC_AV));                                  // :87-92
}

// ---------------------------------------------------------------------------
// Finito adaptive  (Finito_adaptive.jl)
// ---------------------------------------------------------------------------
// Julia's sum(1 ./ γ): pairwise above 1024 terms (Base.mapreduce_impl), restated on the host like the shim does for the
// other variants, so γ̂ carries the reference's rounding
static int prox_vec(ciao_ctx *c, int src, int dst, double gamma);
static double pairwise_recip_sum(const double *v, int64_t lo, int64_t hi) {
    if (hi - lo < 1024) {
        double s = 1 / v[lo];
        for (int64_t i = lo + 1; i < hi; ++i) s += 1 / v[i];
        return s;
    }
    const int64_t mid = lo + ((hi - lo) >> 1);
    return pairwise_recip_sum(v, lo, mid) + pairwise_recip_sum(v, mid, hi);
}

// :59-99.  `perturb` (may be NULL) serves the random restart of the stepsize estimate (:77-83): for a component with
// ∇f_i(x0 + 1) == ∇f_i(x0) it is called as perturb(user, i (1-based), t, xeps) and must fill xeps = x0 .+ rand(t·[−1, 1], size(x0))
// from the host's RNG — components in ascending order, t = 1, 2, 4, … per component, exactly the reference's draw order.
extern "C" int ciao_finito_adaptive_init_cb(ciao_ctx *c, const double *x0, double alpha, double tol_b, ciao_perturb_fn perturb,
                                            void *user) {
    CIAO_TRY(need_rows(c, "ciao_finito_adaptive_init", true));
    if (!x0 || !(alpha > 0) || !(tol_b > 0)) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_finito_adaptive_init: x0 is null, α ≤ 0 or tol_b ≤ 0 (Finito.jl:57-60)");
    if (c->peers.n > 1) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "adaptive Finito keeps N×d tables: not available on row-sharded problems");
    if (use_block_kernel(c) && !c->force_block) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "adaptive Finito is not available for M×d block components / complex data");
    c->algo = ALG_FINITO_ADAPTIVE; c->adapt_alpha = alpha; c->adapt_tol_b = tol_b; c->adapt_backtracks = 0;
    c->cz_valid = false; c->cz_local_valid = false;
    CIAO_TRY(reserve_for_solver(c));
    CIAO_TRY(alloc_table(c));
    const int64_t N = c->N_total;
    if (!c->adapt) CUDA_TRY(cudaMalloc(&c->adapt, (size_t)8 * N * 4 * sizeof(double)));   // one copy per CTA of the cluster
    if (!c->adapt_scal) CUDA_TRY(cudaMalloc(&c->adapt_scal, 8 * sizeof(double)));
    if (!c->adapt_counters) CUDA_TRY(cudaMalloc(&c->adapt_counters, 4 * sizeof(int64_t)));
    CIAO_TRY(upload_vec(c, CIAO_VEC_X0, x0));
    CIAO_TRY(run_pass_finish(c, PASS_GRAD, ctx_vec(c, CIAO_VEC_X0), false, nullptr, 1.0, 1.0, ctx_vec(c, CIAO_VEC_TMP)));  // sum(∇f)  :90
    CIAO_TRY(run_adaptive_init(c, ctx_vec(c, CIAO_VEC_X0), alpha));                 // :65-87 with t = 1
    std::vector<double> gam((size_t)N);
    CUDA_TRY(cudaMemcpy2DAsync(gam.data(), sizeof(double), c->adapt, 4 * sizeof(double), sizeof(double), (size_t)N, cudaMemcpyDeviceToHost,
                               c->stream));
    int h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, c->err_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h) {   // some ∇f_i(x0 + 1) == ∇f_i(x0): the random restart, :77-83
        CUDA_TRY(cudaMemsetAsync(c->err_dev, 0, sizeof(int), c->stream));
        if (!perturb) {
            c->algo = 0;
            CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "adaptive Finito: ∇f_i(x0 + 1) == ∇f_i(x0) for some i; the reference then perturbs x0 at random "
                      "(Finito_adaptive.jl:77-83) — pass a perturbation callback (ciao_finito_adaptive_init_cb) or choose another x0");
        }
        std::vector<double> nmg((size_t)N), xeps((size_t)c->d);
        CUDA_TRY(cudaMemcpy2D(nmg.data(), sizeof(double), c->adapt + 3, 4 * sizeof(double), sizeof(double), (size_t)N, cudaMemcpyDeviceToHost));
        const double sqrt_d = sqrt((double)c->d);
        for (int64_t i = 0; i < N; ++i) {
            if (!(nmg[(size_t)i] < 2.220446049250313e-16)) continue;
            double nm = nmg[(size_t)i];
            int64_t t = 1;                                                          // :76
            while (nm < 2.220446049250313e-16) {                                    // :77
                if (t > ((int64_t)1 << 52) || perturb(user, i + 1, t, xeps.data()) != 0) {   // :79 (the host's RNG)
                    c->algo = 0;
                    CIAO_FAIL(CIAO_ERR_INVALID, "adaptive Finito: the perturbation callback failed for component %lld (t = %lld)",
                              (long long)(i + 1), (long long)t);
                }
                CIAO_TRY(upload_vec(c, CIAO_VEC_X, xeps.data()));
                CIAO_TRY(run_adaptive_retry(c, i, ctx_vec(c, CIAO_VEC_X0), ctx_vec(c, CIAO_VEC_X), c->partial + c->d_pad + 2));   // :80-81
                CUDA_TRY(cudaMemcpyAsync(&nm, c->partial + c->d_pad + 2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                CUDA_TRY(cudaStreamSynchronize(c->stream));
                t *= 2;                                                             // :82
            }
            double L_int = nm / ((double)t * sqrt_d);                               // :84
            L_int /= (double)N;                                                     // :85
            gam[(size_t)i] = alpha / L_int;                                         // :86
            CUDA_TRY(cudaMemcpyAsync(c->adapt + 4 * i, &gam[(size_t)i], sizeof(double), cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
    }
    int chunks = 0;
    CIAO_TRY(run_adaptive_sdivg(c, ctx_vec(c, CIAO_VEC_X0), &chunks));              // partial sums of x0 ./ γ_i
    launch_reduce_ws(c, c->ws, c->ws + (size_t)chunks * c->d_pad, chunks, c->partial, c->partial + c->d_pad, 0, 1);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    for (int r = 1; r < 8; ++r)   // the per-CTA copies of {γ_i, f_i, c_i} start identical
        CUDA_TRY(cudaMemcpyAsync(c->adapt + (size_t)r * 4 * N, c->adapt, (size_t)N * 4 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->hat_gamma = 1 / pairwise_recip_sum(gam.data(), 0, N);                        // :89
    CUDA_TRY(cudaMemcpyAsync(c->adapt_scal, &c->hat_gamma, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CIAO_TRY(run_adaptive_av(c, c->partial, ctx_vec(c, CIAO_VEC_TMP), c->hat_gamma));  // :90
    CIAO_TRY(prox_vec(c, CIAO_VEC_AV, CIAO_VEC_Z, c->hat_gamma));                   // :91
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CIAO_OK;
}

extern "C" int ciao_finito_adaptive_init(ciao_ctx *c, const double *x0, double alpha, double tol_b) {
    return ciao_finito_adaptive_init_cb(c, x0, alpha, tol_b, nullptr, nullptr);
}

extern "C" int ciao_finito_adaptive_steps(ciao_ctx *c, const int64_t *idx, int64_t K, int64_t *steps_done) {
    CIAO_TRY(need_rows(c, "ciao_finito_adaptive_steps", true));
    if (c->algo != ALG_FINITO_ADAPTIVE) CIAO_FAIL(CIAO_ERR_STATE, "ciao_finito_adaptive_steps before ciao_finito_adaptive_init");
    if (K < 0) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_finito_adaptive_steps: negative step count");
    if (steps_done) *steps_done = 0;
    if (K == 0) return CIAO_OK;
    const int64_t *raw;
    CIAO_TRY(fetch_raw_indices(c, idx, K, &raw));
    CIAO_TRY(launch_prep_indices(c, raw, K, c->N_total, c->idx_prep));
    CIAO_TRY(check_err_flag(c));   // an out-of-range index must not reach the table updates
    CIAO_TRY(run_seq_adaptive(c, c->idx_prep, K, c->adapt_alpha, c->adapt_tol_b));  // :101-160
    int64_t cnt[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(cnt, c->adapt_counters, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(&c->hat_gamma, c->adapt_scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->adapt_backtracks += cnt[1];
    if (steps_done) *steps_done = cnt[0];
    return CIAO_OK;
}

// state.γ, state.fi_x, c_i with ∇f_i(x_i) = c_i·a_i (each N entries, any may be NULL), state.hat_γ, total reductions of γ
extern "C" int ciao_finito_adaptive_get(ciao_ctx *c, double *gamma_N, double *fi_x_N, double *coef_N, double *hat_gamma, int64_t *backtracks) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_finito_adaptive_get: null context");
    if (c->algo != ALG_FINITO_ADAPTIVE || !c->adapt) CIAO_FAIL(CIAO_ERR_STATE, "ciao_finito_adaptive_get before ciao_finito_adaptive_init");
    CUDA_TRY(cudaSetDevice(c->device));
    double *outs[3] = {gamma_N, fi_x_N, coef_N};
    for (int j = 0; j < 3; ++j)
        if (outs[j])
            CUDA_TRY(cudaMemcpy2DAsync(outs[j], sizeof(double), c->adapt + j, 4 * sizeof(double), sizeof(double), (size_t)c->N_total,
                                       cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (hat_gamma) *hat_gamma = c->hat_gamma;
    if (backtracks) *backtracks = c->adapt_backtracks;
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
// SAGA / SAG
// ---------------------------------------------------------------------------
extern "C" int ciao_saga_init(ciao_ctx *c, const double *x0, double gamma, int sag) {
    CIAO_TRY(need_rows(c, "ciao_saga_init", false));   // shards: table init pass + allreduce of Σ s_i
    if (!x0 || !(gamma > 0)) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_saga_init: x0 is null or γ ≤ 0 (SAGA.jl:37)");
    c->algo = ALG_SAGA; c->gamma = gamma; c->sag = sag ? 1 : 0;
    CIAO_TRY(reserve_for_solver(c));
    CIAO_TRY(alloc_table(c));
    CIAO_TRY(upload_vec(c, CIAO_VEC_X0, x0));
    CIAO_TRY(run_pass_finish(c, PASS_SAGA_INIT, ctx_vec(c, CIAO_VEC_X0), false, nullptr, 1.0, (double)c->N_total,
                             ctx_vec(c, CIAO_VEC_AV)));                              // :41-47
    saga_z0_kernel<<<blocks_for(c->d_pad), 256, 0, c->stream>>>(ctx_vec(c, CIAO_VEC_X0), ctx_vec(c, CIAO_VEC_Z), c->d_pad, gamma, c->reg);  // :48
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

extern "C" int ciao_saga_steps(ciao_ctx *c, const int64_t *idx, int64_t K) {
    CIAO_TRY(need_rows(c, "ciao_saga_steps", true));
    if (c->algo != ALG_SAGA) CIAO_FAIL(CIAO_ERR_STATE, "ciao_saga_steps before ciao_saga_init");
    CIAO_TRY(need_whole_table(c, "ciao_saga_steps"));
    if (K == 0) return CIAO_OK;
    const int64_t *raw;
    CIAO_TRY(fetch_raw_indices(c, idx, K, &raw));
    CIAO_TRY(launch_prep_indices(c, raw, K, c->N_total, c->idx_prep));
    return run_seq(c, ALG_SAGA, c->idx_prep, K, 1.0);
}

// ---------------------------------------------------------------------------
// Finito / LFinito
// ---------------------------------------------------------------------------
static int prox_vec(ciao_ctx *c, int src, int dst, double gamma) {
    prox_vec_kernel<<<blocks_for(c->d_pad), 256, 0, c->stream>>>(ctx_vec(c, src), ctx_vec(c, dst), c->d_pad, gamma, c->reg);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

extern "C" int ciao_finito_init(ciao_ctx *c, const double *x0, const double *gamma_N, double hat_gamma) {
    CIAO_TRY(need_rows(c, "ciao_finito_init", false));  // shards: table init pass + allreduce of Σ s_i/γ_i
    if (!x0 || !gamma_N || !(hat_gamma > 0)) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_finito_init: null argument or γ̂ ≤ 0");
    c->algo = ALG_FINITO; c->hat_gamma = hat_gamma;
    CIAO_TRY(reserve_for_solver(c));
    CIAO_TRY(alloc_table(c));
    CIAO_TRY(set_gammas(c, gamma_N, true));
    CIAO_TRY(upload_vec(c, CIAO_VEC_X0, x0));
    CIAO_TRY(run_pass_finish(c, PASS_FINITO_INIT, ctx_vec(c, CIAO_VEC_X0), false, nullptr, hat_gamma, 1.0,
                             ctx_vec(c, CIAO_VEC_AV)));                              // :76-83
    return prox_vec(c, CIAO_VEC_AV, CIAO_VEC_Z, hat_gamma);                        // :84
}

static int batched_indices(ciao_ctx *c, const int64_t *idx, const int64_t *batch_ptr, int64_t n_batches, int64_t *n_idx_out,
                           const int64_t **ptr_dev_out = nullptr) {
    if (!batch_ptr || n_batches < 0) CIAO_FAIL(CIAO_ERR_INVALID, "batch_ptr is null or n_batches < 0");
    int64_t ends[2] = {0, 0};
    CUDA_TRY(cudaMemcpy(&ends[0], batch_ptr, sizeof(int64_t), cudaMemcpyDefault));
    CUDA_TRY(cudaMemcpy(&ends[1], batch_ptr + n_batches, sizeof(int64_t), cudaMemcpyDefault));
    if (ends[0] != 0 || ends[1] < 0) CIAO_FAIL(CIAO_ERR_INVALID, "batch_ptr must start at 0 and be non-decreasing");
    const int64_t n_idx = ends[1];
    *n_idx_out = n_idx;
    if (n_idx == 0) return CIAO_OK;
    const int64_t *raw, *ptr_dev;
    CIAO_TRY(fetch_raw_indices(c, idx, n_idx, &raw));
    CIAO_TRY(upload_ptr(c, batch_ptr, n_batches + 1, &ptr_dev));
    if (ptr_dev_out) *ptr_dev_out = ptr_dev;
    CIAO_TRY(launch_prep_indices(c, raw, n_idx, c->N_total, c->idx_prep));
    return launch_mark_batch_ends(c, ptr_dev, n_batches, n_idx, c->idx_prep);
}

extern "C" int ciao_finito_steps(ciao_ctx *c, const int64_t *idx, const int64_t *batch_ptr, int64_t n_batches) {
    CIAO_TRY(need_rows(c, "ciao_finito_steps", false));
    if (c->algo != ALG_FINITO) CIAO_FAIL(CIAO_ERR_STATE, "ciao_finito_steps before ciao_finito_init");
    // Row shards (one process per GPU): static minibatches shard by row owner — every rank streams its part of a batch, the
    // batch's Σ is all-reduced in the tail kernel (collective: all ranks make the same call).  Single samples / scattered
    // batches need the whole table on one GPU.
    const bool sharded = c->n_rows != c->N_total;
    // static minibatches (contiguous rows, Finito_basic.jl:52-57) of ≥ BATCH_MIN_ROWS rows: one streaming pass per batch
    if (idx && batch_ptr && n_batches > 0 && !is_device_ptr(idx) && !is_device_ptr(batch_ptr) && !use_block_kernel(c)) {
        bool contiguous = true;
        int64_t longest = 0;
        for (int64_t j = 0; j < n_batches && contiguous; ++j) {
            const int64_t lo = batch_ptr[j], hi = batch_ptr[j + 1];
            if (lo < 0 || hi < lo) CIAO_FAIL(CIAO_ERR_INVALID, "batch_ptr must be non-decreasing");
            longest = std::max(longest, hi - lo);
            if (hi > lo && (idx[lo] < 1 || idx[lo] + (hi - lo) - 1 > c->N_total)) contiguous = false;
            for (int64_t t = lo + 1; t < hi && contiguous; ++t) contiguous = idx[t] == idx[t - 1] + 1;
        }
        if (contiguous && batch_ptr[0] == 0 && longest >= BATCH_MIN_ROWS) {
            std::vector<int64_t> win;   // [first rows…, lengths…] of the non-empty batches
            for (int64_t j = 0; j < n_batches; ++j)
                if (batch_ptr[j + 1] > batch_ptr[j]) win.push_back(idx[batch_ptr[j]] - 1);
            const int64_t nbw = (int64_t)win.size();
            for (int64_t j = 0; j < n_batches; ++j)
                if (batch_ptr[j + 1] > batch_ptr[j]) win.push_back(batch_ptr[j + 1] - batch_ptr[j]);
            CUDA_TRY(cudaEventRecord(c->ev_sa, c->stream));
            int rc = CIAO_ERR_UNSUPPORTED;
            if ((!sharded || (c->p2p_ready && c->world > 1)) && c->batch_persistent && nbw > 1 && nbw < ((int64_t)1 << 22)) {   // all batches in one cooperative launch
                std::vector<int64_t> loc(win);   // row shards: the part of every window this context holds, in local row numbers
                if (sharded)
                    for (int64_t j = 0; j < nbw; ++j) CIAO_TRY(local_window(c, win[j], win[nbw + j], &loc[j], &loc[nbw + j]));
                const int64_t *win_dev;
                CIAO_TRY(upload_ptr(c, loc.data(), 2 * nbw, &win_dev));
                // do rows repeat inside the call, and if so always in the same window?  (sorted by first row: equal or disjoint neighbours)
                std::vector<std::pair<int64_t, int64_t>> ws_sorted((size_t)nbw);
                for (int64_t j = 0; j < nbw; ++j) ws_sorted[(size_t)j] = {win[j], win[nbw + j]};
                std::sort(ws_sorted.begin(), ws_sorted.end());
                int windows = BATCH_WINDOWS_DISJOINT;
                for (int64_t j = 1; j < nbw && windows != BATCH_WINDOWS_ANY; ++j) {
                    const auto &u = ws_sorted[(size_t)j - 1], &v = ws_sorted[(size_t)j];
                    if (u == v) windows = BATCH_WINDOWS_ALIGNED;
                    else if (u.first + u.second > v.first) windows = BATCH_WINDOWS_ANY;
                }
                rc = run_batch_sequence(c, BATCH_FINITO, win_dev, win_dev + nbw, nbw, longest, windows, sharded);
                if (rc != CIAO_OK && rc != CIAO_ERR_UNSUPPORTED) return rc;
            }
            if (rc == CIAO_ERR_UNSUPPORTED)
                for (int64_t j = 0; j < nbw; ++j) CIAO_TRY(run_batch_step(c, BATCH_FINITO, win[j], win[nbw + j]));
            CUDA_TRY(cudaEventRecord(c->ev_sb, c->stream));
            c->timing.last_seq_steps = batch_ptr[n_batches];
            c->seq_timed = true;
            return CIAO_OK;
        }
    }
    CIAO_TRY(need_whole_table(c, "ciao_finito_steps (single samples or scattered batches)"));
    int64_t n_idx = 0;
    CIAO_TRY(batched_indices(c, idx, batch_ptr, n_batches, &n_idx));
    return run_seq(c, ALG_FINITO, c->idx_prep, n_idx, 1.0);
}

extern "C" int ciao_lfinito_init(ciao_ctx *c, const double *x0, const double *gamma_N, double hat_gamma) {
    CIAO_TRY(need_rows(c, "ciao_lfinito_init", false));   // the init is a pass: it shards like the full gradient
    if (!x0 || !gamma_N || !(hat_gamma > 0)) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_lfinito_init: null argument or γ̂ ≤ 0");
    c->algo = ALG_LFINITO; c->hat_gamma = hat_gamma;
    c->cz_valid = false; c->cz_local_valid = false;
    CIAO_TRY(reserve_for_solver(c));
    CIAO_TRY(set_gammas(c, gamma_N, true));
    CIAO_TRY(upload_vec(c, CIAO_VEC_X0, x0));
    CIAO_TRY(run_pass_finish(c, PASS_GRAD, ctx_vec(c, CIAO_VEC_X0), false, ctx_vec(c, CIAO_VEC_X0), -(hat_gamma / (double)c->N_total), 1.0,
                             ctx_vec(c, CIAO_VEC_AV)));                              // :68-72
    CIAO_TRY(copy_vec(c, CIAO_VEC_Z, CIAO_VEC_AV));                                // placeholders, ctor :33-35
    return copy_vec(c, CIAO_VEC_Z_FULL, CIAO_VEC_AV);
}

extern "C" int ciao_lfinito_outer(ciao_ctx *c, const int64_t *batch_order, int64_t n_batches, int64_t r) {
    // row shards without attached peers: only the minibatch sweep (r ≥ BATCH_MIN_ROWS) shards — by row owner, one exchange per batch
    CIAO_TRY(need_rows(c, "ciao_lfinito_outer", !(r >= BATCH_MIN_ROWS && c && c->n_rows != c->N_total && c->world > 1)));
    if (c->algo != ALG_LFINITO) CIAO_FAIL(CIAO_ERR_STATE, "ciao_lfinito_outer before ciao_lfinito_init");
    if (!batch_order || r <= 0 || n_batches < 0) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_lfinito_outer: bad arguments");
    const int64_t N = c->N_total, nb = (N + r - 1) / r;
    // where does the (possibly short) last static batch sit in the order?
    std::vector<int64_t> order((size_t)n_batches);
    CUDA_TRY(cudaMemcpy(order.data(), batch_order, (size_t)n_batches * sizeof(int64_t), cudaMemcpyDefault));
    const int64_t last_len = N - r * (nb - 1);
    int64_t short_pos = -1, total = 0;
    for (int64_t jj = 0; jj < n_batches; ++jj) {
        const int64_t j = order[jj];
        if (j < 1 || j > nb) CIAO_FAIL(CIAO_ERR_INVALID, "batch_order[%lld] = %lld not in 1..%lld", (long long)jj, (long long)j, (long long)nb);
        if (j == nb && last_len != r) {
            if (short_pos >= 0) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "the short last batch may appear once per sweep");
            short_pos = jj;
        }
        total += (j == nb) ? last_len : r;
    }
    CIAO_TRY(prox_vec(c, CIAO_VEC_AV, CIAO_VEC_Z_FULL, c->hat_gamma));             // :83
    CIAO_TRY(run_pass_finish(c, PASS_GRAD, ctx_vec(c, CIAO_VEC_Z_FULL), c->cache_cz, ctx_vec(c, CIAO_VEC_Z_FULL),
                             -(c->hat_gamma / (double)N), 1.0, ctx_vec(c, CIAO_VEC_AV)));  // :85-88
    if (total == 0) return CIAO_OK;
    const bool sharded = c->n_rows != c->N_total;
    if (r >= BATCH_MIN_ROWS && (!sharded || c->world > 1) && !use_block_kernel(c)) {  // minibatch sweep: prox + one streaming pass per batch (:91-100)
        CUDA_TRY(cudaEventRecord(c->ev_sa, c->stream));
        int rc = CIAO_ERR_UNSUPPORTED;
        if ((!sharded || c->p2p_ready) && c->batch_persistent && n_batches > 1 && n_batches < ((int64_t)1 << 22)) {  // the whole sweep in one cooperative launch
            std::vector<int64_t> win((size_t)2 * n_batches);
            for (int64_t jj = 0; jj < n_batches; ++jj) {
                win[jj] = r * (order[jj] - 1);
                win[n_batches + jj] = (order[jj] == nb) ? last_len : r;
                if (sharded) {   // the part of the window this context holds, in local row numbers
                    const int64_t glo = win[jj], gn = win[n_batches + jj];
                    CIAO_TRY(local_window(c, glo, gn, &win[jj], &win[n_batches + jj]));
                }
            }
            const int64_t *win_dev;
            CIAO_TRY(upload_ptr(c, win.data(), 2 * n_batches, &win_dev));
            CIAO_TRY(prox_vec(c, CIAO_VEC_AV, CIAO_VEC_Z, c->hat_gamma));          // :92 of the first batch; later ones in the kernel
            rc = run_batch_sequence(c, BATCH_LFINITO, win_dev, win_dev + n_batches, n_batches, r, BATCH_WINDOWS_ANY, sharded);   // no table: the window kind is irrelevant
            if (rc != CIAO_OK && rc != CIAO_ERR_UNSUPPORTED) return rc;
        }
        if (rc == CIAO_ERR_UNSUPPORTED)
            for (int64_t jj = 0; jj < n_batches; ++jj) {   // z = prox_g(av) (:92) of batch jj + 1 is formed by the tail kernel of batch jj
                const int64_t j = order[jj];
                if (jj == 0) CIAO_TRY(prox_vec(c, CIAO_VEC_AV, CIAO_VEC_Z, c->hat_gamma));
                CIAO_TRY(run_batch_step(c, BATCH_LFINITO, r * (j - 1), (j == nb) ? last_len : r, jj + 1 == n_batches));
            }
        CUDA_TRY(cudaEventRecord(c->ev_sb, c->stream));
        c->timing.last_seq_steps = total;
        c->seq_timed = true;
        return CIAO_OK;
    }
    CIAO_TRY(reserve_idx(c, (size_t)total));
    const int64_t *order_dev;
    CIAO_TRY(upload_ptr(c, order.data(), n_batches, &order_dev));
    CIAO_TRY(launch_expand_batches(c, order_dev, n_batches, r, N, nb, short_pos, c->idx_prep));
    return run_seq(c, ALG_LFINITO, c->idx_prep, total, 1.0);                       // :91-100
}

// ---------------------------------------------------------------------------
// ProShI
// ---------------------------------------------------------------------------
static int need_blocks(ciao_ctx *c, const char *who) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "%s: null context", who);
    if (c->loss_kind != CIAO_LOSS_DIAGQUAD) CIAO_FAIL(CIAO_ERR_STATE, "%s: no sharing problem set (ciao_set_blocks first)", who);
    CUDA_TRY(cudaSetDevice(c->device));
    return CIAO_OK;
}

extern "C" int ciao_proshi_init(ciao_ctx *c, const double *x0, const double *gamma_N, double hat_gamma) {
    CIAO_TRY(need_blocks(c, "ciao_proshi_init"));
    if (!x0 || !gamma_N || !(hat_gamma > 0)) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_proshi_init: null argument or γ̂ ≤ 0");
    c->algo = ALG_PROSHI; c->hat_gamma = hat_gamma;
    CIAO_TRY(reserve_for_solver(c));
    CIAO_TRY(alloc_table(c));
    CIAO_TRY(set_gammas(c, gamma_N, false));
    CIAO_TRY(upload_vec(c, CIAO_VEC_X0, x0));
    CIAO_TRY(run_proshi_init(c, ctx_vec(c, CIAO_VEC_X0)));                         // :76-80
    CIAO_TRY(run_finish(c, nullptr, 1.0, 1.0, ctx_vec(c, CIAO_VEC_AV)));           // :83
    return run_proshi_dual(c);                                                     // :84-86
}

extern "C" int ciao_proshi_steps(ciao_ctx *c, const int64_t *idx, const int64_t *batch_ptr, int64_t n_batches) {
    CIAO_TRY(need_blocks(c, "ciao_proshi_steps"));
    if (c->algo != ALG_PROSHI) CIAO_FAIL(CIAO_ERR_STATE, "ciao_proshi_steps before ciao_proshi_init");
    int64_t n_idx = 0;
    const int64_t *ptr_dev = nullptr;
    CIAO_TRY(batched_indices(c, idx, batch_ptr, n_batches, &n_idx, &ptr_dev));
    return run_proshi_steps(c, c->idx_prep, n_idx, ptr_dev, n_batches);
}

extern "C" int ciao_proshi_solution(ciao_ctx *c, double *S_out) {
    CIAO_TRY(need_blocks(c, "ciao_proshi_solution"));
    if (c->algo != ALG_PROSHI) CIAO_FAIL(CIAO_ERR_STATE, "ciao_proshi_solution before ciao_proshi_init");
    CIAO_TRY(run_proshi_solution(c));
    if (S_out) {
        CUDA_TRY(cudaMemcpy2DAsync(S_out, (size_t)c->d * 8, c->table, (size_t)c->d_pad * 8, (size_t)c->d * 8, (size_t)c->N_total,
                                   cudaMemcpyDefault, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
// state access
// ---------------------------------------------------------------------------
extern "C" int ciao_get_vec(ciao_ctx *c, int which, double *out, int64_t len) {
    if (!c || !out) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_get_vec: null argument");
    if (!c->vecs) CIAO_FAIL(CIAO_ERR_STATE, "ciao_get_vec: no problem set");
    if (which < 0 || which > CIAO_VEC_X || len != c->d) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_get_vec: bad vector id or length (d = %lld)", (long long)c->d);
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(out, ctx_vec(c, which), (size_t)len * sizeof(double), cudaMemcpyDefault, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_err_flag(c);
}

extern "C" int ciao_set_vec(ciao_ctx *c, int which, const double *in, int64_t len) {
    if (!c || !in) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_vec: null argument");
    if (!c->vecs) CIAO_FAIL(CIAO_ERR_STATE, "ciao_set_vec: no problem set");
    if (which < 0 || which > CIAO_VEC_X || len != c->d) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_vec: bad vector id or length");
    CUDA_TRY(cudaSetDevice(c->device));
    c->cz_valid = false; c->cz_local_valid = false;  // z_full may have changed: the cached c_i(z_full) no longer apply
    return upload_vec(c, which, in);
}

extern "C" int ciao_get_table_rows(ciao_ctx *c, int64_t i0, int64_t n, double *out) {
    if (!c || !out) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_get_table_rows: null argument");
    if (!c->table) CIAO_FAIL(CIAO_ERR_STATE, "ciao_get_table_rows: no table (SAGA/Finito/ProShI init first)");
    if (i0 < 0 || n < 0 || i0 + n > c->n_rows) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_get_table_rows: rows out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    if (n == 0) return CIAO_OK;
    CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)c->d * 8, c->table + i0 * c->d_pad, (size_t)c->d_pad * 8, (size_t)c->d * 8, (size_t)n,
                               cudaMemcpyDefault, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CIAO_OK;
}

// rows [i0, i0+n) of the table ← host/device array (n×d row-major): the restore half of "the iterator state is the checkpoint"
extern "C" int ciao_set_table_rows(ciao_ctx *c, int64_t i0, int64_t n, const double *in) {
    if (!c || !in) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_table_rows: null argument");
    if (!c->table) CIAO_FAIL(CIAO_ERR_STATE, "ciao_set_table_rows: no table (a table solver's init or ciao_solver_restore first)");
    if (i0 < 0 || n < 0 || i0 + n > c->n_rows) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_set_table_rows: rows out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    if (n == 0) return CIAO_OK;
    CUDA_TRY(cudaMemcpy2DAsync(c->table + i0 * c->d_pad, (size_t)c->d_pad * 8, in, (size_t)c->d * 8, (size_t)c->d * 8, (size_t)n,
                               cudaMemcpyDefault, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CIAO_OK;
}

// Re-creates a solver's device state WITHOUT running its init pass, for a restore from a checkpoint: sets the algorithm, its
// stepsizes (γ or γ_i, γ̂) and flags, allocates the table; the caller then restores the vectors with ciao_set_vec (z, z_full,
// w, av) and the table with ciao_set_table_rows and continues with the *_steps / *_epoch calls.
//   algo: 1 SVRG (gamma, flag = plus) | 2 SAGA (gamma, flag = sag) | 3 Finito | 4 LFinito | 5 ProShI (gamma_N, hat_gamma)
extern "C" int ciao_solver_restore(ciao_ctx *c, int algo, double gamma, int flag, const double *gamma_N, double hat_gamma) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_solver_restore: null context");
    if (algo == ALG_PROSHI) CIAO_TRY(need_blocks(c, "ciao_solver_restore"));
    else CIAO_TRY(need_rows(c, "ciao_solver_restore", false));
    switch (algo) {
        case ALG_SVRG:
        case ALG_SAGA:
            if (!(gamma > 0)) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_solver_restore: γ ≤ 0");
            c->gamma = gamma;
            if (algo == ALG_SVRG) c->plus = flag ? 1 : 0;
            else c->sag = flag ? 1 : 0;
            break;
        case ALG_FINITO:
        case ALG_LFINITO:
        case ALG_PROSHI:
            if (!gamma_N || !(hat_gamma > 0)) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_solver_restore: γ_i missing or γ̂ ≤ 0");
            c->hat_gamma = hat_gamma;
            break;
        default:
            CIAO_FAIL(CIAO_ERR_INVALID, "ciao_solver_restore: algo must be 1 (SVRG) … 5 (ProShI)");
    }
    c->algo = algo;
    c->cz_valid = false; c->cz_local_valid = false;
    CIAO_TRY(reserve_for_solver(c));
    if (algo == ALG_SAGA || algo == ALG_FINITO || algo == ALG_PROSHI) CIAO_TRY(alloc_table(c));
    if (algo == ALG_FINITO || algo == ALG_LFINITO) CIAO_TRY(set_gammas(c, gamma_N, true));
    if (algo == ALG_PROSHI) {
        CIAO_TRY(set_gammas(c, gamma_N, false));
        CIAO_TRY(run_proshi_gpair(c));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CIAO_OK;
}

extern "C" int ciao_table_colsum(ciao_ctx *c, double *out) {
    if (!c || !out) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_table_colsum: null argument");
    if (!c->table) CIAO_FAIL(CIAO_ERR_STATE, "ciao_table_colsum: no table");
    CUDA_TRY(cudaSetDevice(c->device));
    CIAO_TRY(run_table_colsum(c, c->n_rows));
    CUDA_TRY(cudaMemcpyAsync(out, c->partial, (size_t)c->d * sizeof(double), cudaMemcpyDefault, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------
extern "C" int ciao_stage_indices(ciao_ctx *c, const int64_t *idx_host, int64_t n) {
    if (!c || !idx_host || n <= 0) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_stage_indices: bad arguments");
    CUDA_TRY(cudaSetDevice(c->device));
    c->staged = 0;
    CIAO_TRY(reserve_idx(c, (size_t)n));
    CUDA_TRY(cudaMemcpyAsync(c->idx_raw, idx_host, (size_t)n * sizeof(int64_t), cudaMemcpyDefault, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->staged = n;
    return CIAO_OK;
}

extern "C" int ciao_timer_begin(ciao_ctx *c) {
    if (!c) CIAO_FAIL(CIAO_ERR_INVALID, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->tm_a, c->stream));
    return CIAO_OK;
}

extern "C" int ciao_timer_end(ciao_ctx *c, float *ms) {
    if (!c || !ms) CIAO_FAIL(CIAO_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->tm_b, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->tm_b));
    CUDA_TRY(cudaEventElapsedTime(ms, c->tm_a, c->tm_b));
    return CIAO_OK;
}

extern "C" int ciao_last_seq_placement(ciao_ctx *c, int *smid16, int *n_ctas) {
    if (!c || !smid16 || !n_ctas) CIAO_FAIL(CIAO_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(smid16, c->seq_smid, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *n_ctas = c->seq_smid_n;
    return CIAO_OK;
}

// SM cycles and wall nanoseconds CTA 0 of the last sequential cluster kernel spent in it: cycles/ns = the SM clock (GHz) the kernel
// actually ran at — a latency-bound kernel's µs/step scales with it, and NVML's 100 ms samples do not resolve short dips
extern "C" int ciao_last_seq_clock(ciao_ctx *c, int64_t *cycles, int64_t *ns) {
    if (!c || !cycles || !ns) CIAO_FAIL(CIAO_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    long long h[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(h, c->seq_smid + 16, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *cycles = h[0];
    *ns = h[1];
    return CIAO_OK;
}

extern "C" int ciao_last_timing(ciao_ctx *c, ciao_timing *out) {
    if (!c || !out) CIAO_FAIL(CIAO_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->pass_timed) CUDA_TRY(cudaEventElapsedTime(&c->timing.last_pass_ms, c->ev_pa, c->ev_pb));
    if (c->pass_timed && c->tail_timed) CUDA_TRY(cudaEventElapsedTime(&c->timing.last_tail_ms, c->ev_pb, c->ev_pc));
    if (c->seq_timed) CUDA_TRY(cudaEventElapsedTime(&c->timing.last_seq_ms, c->ev_sa, c->ev_sb));
    *out = c->timing;
    return CIAO_OK;
}

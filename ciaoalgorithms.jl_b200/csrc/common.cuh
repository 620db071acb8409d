// common.cuh — context, error plumbing, sm_100a PTX wrappers (mbarrier, TMA bulk
// copy, cluster/DSMEM) and the operator device functions shared by all kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ciao_cuda.h"
// NVTX ranges around every pass / persistent-kernel call (header-only nvtx3: a no-op unless a profiler is attached)
#include <nvtx3/nvToolsExt.h>
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
#ifdef __CUDACC__
#include "fastmath.cuh"
#endif

// ---------------------------------------------------------------------------
// HBM data layout
//   records : [n_rows][ld]  fp64, ld = d_pad + 8, d_pad = round_up(d, 4)
//             row record i = [ a_i (d, zero padded to d_pad) | tail of 8 scalars ]
//             tail = b_i or y_i | λ_i or μ_i | γ_i | γ_i/N | γ̂/γ_i | 0 | 0 | 0
//             → one TMA bulk copy brings a row and its scalars (no per-step division, no
//             dependent scalar loads); records are 32-byte aligned.
//   table   : [N][d_pad]    fp64  (SAGA gradients / Finito, ProShI s_i)
//   vecs    : [CIAO_NUM_VECS][d_pad]  state vectors
// ---------------------------------------------------------------------------
#define CIAO_TAIL 8
#define TAIL_B 0       // b_i (LS) or y_i (logistic)
#define TAIL_LAM 1     // λ_i or μ_i
#define TAIL_GAM 2     // γ_i                      (Finito/LFinito, Finito_basic.jl:61-74)
#define TAIL_GAM_N 3   // γ_i / N                  (Finito_basic.jl:79,113)
#define TAIL_HAT_GAM 4 // γ̂ / γ_i                  (Finito_basic.jl:115, Finito_LFinito.jl:98)
#define CIAO_TAIL_USED 6  // doubles of the tail the sequential kernels stage (16-byte multiple)
#define CIAO_NUM_VECS 8
#define CIAO_VEC_X0 6
#define CIAO_VEC_TMP 7

#define CIAO_IDX_MASK 0x0000FFFFFFFFFFFFll
// prep_indices_kernel flags a step whose row/block index also occurs 1 … CIAO_HAZARD_WINDOW − 1 steps earlier; the sequential
// kernels stage table rows at most that many steps ahead (seq_impl.cuh SEQ_D = 8, proshi.cu PROSHI_D = 16)
#define CIAO_HAZARD_WINDOW 20
#define CIAO_HAZ_DIST_SHIFT 48         // bits 48..52 of a flagged step: distance (1 … CIAO_HAZARD_WINDOW − 1) to the previous occurrence
#define CIAO_FLAG_HAZARD (1ll << 62)  // same row was written < prefetch-depth steps ago: reload the table row
#define CIAO_FLAG_PROX (1ll << 61)    // batch boundary: apply prox_g at this step (after: Finito/ProShI, before: LFinito)

enum { ALG_SVRG = 1, ALG_SAGA = 2, ALG_FINITO = 3, ALG_LFINITO = 4, ALG_PROSHI = 5, ALG_FINITO_ADAPTIVE = 6 };

// Row shards reachable from this GPU: shard s holds rows [start[s], start[s+1]) at base[s] (its own HBM, or a peer's
// HBM mapped through CUDA IPC and read over NVLink).  n = 1: everything is local.
#define CIAO_MAX_PEERS 8
#define CIAO_MAX_DEVICES 16
struct PeerTable {
    const double *base[CIAO_MAX_PEERS];
    int64_t start[CIAO_MAX_PEERS + 1];
    int n;
};

// Exchange arena: mail[2][CIAO_MAX_PEERS][P2P_CAP] doubles (double-buffered by collective parity: a slot of collective n is
// rewritten by collective n + 2, which a peer can only start after it has seen this rank's flag of n + 1, i.e. after this
// rank has finished reading n), then flags[CIAO_MAX_PEERS][P2P_CHUNKS] sequence numbers (one per source rank and 32-column chunk).
#define P2P_CAP 8224                           // d_pad ≤ 8192 columns + the scalar, padded to a multiple of 32
#define P2P_CHUNKS (P2P_CAP / 32)
#define P2P_MAIL_DOUBLES ((size_t)2 * CIAO_MAX_PEERS * P2P_CAP)
// … then, 256-byte aligned, the flagged-word area of the persistent minibatch kernel on row shards (batch.cu): xw[2][CIAO_MAX_PEERS]
// [P2P_CAP] doubles as two 64-bit words {32 data bits | 32-bit epoch} each, double-buffered by batch parity
#define P2P_LL_OFFSET ((P2P_MAIL_DOUBLES * 8 + (size_t)CIAO_MAX_PEERS * P2P_CHUNKS * 4 + 255) / 256 * 256)
#define P2P_LL_WORDS ((size_t)2 * CIAO_MAX_PEERS * P2P_CAP * 2)
#define P2P_ARENA_BYTES (P2P_LL_OFFSET + P2P_LL_WORDS * 8 + 256)

// interleaved shards: global row of local row r, and the local rows [*lo_loc, *lo_loc + *n_loc) of the global window [lo, lo + n)
__host__ __device__ __forceinline__ int64_t il_global(int64_t r, int64_t B, int rank, int world) {
    return ((r / B) * world + rank) * B + r % B;
}
static inline int64_t il_count(int64_t n, int64_t B, int rank, int world) {   // rows of [0, n) this rank owns
    const int64_t sb = B * world, rem = n % sb - (int64_t)rank * B;
    return (n / sb) * B + (rem < 0 ? 0 : (rem > B ? B : rem));
}

struct RegParams {
    int kind;
    double lambda;       // NormL1
    double lo_s, hi_s;   // IndBox scalar bounds
    const double *lo_v;  // IndBox vector bounds (device, length d_pad) or nullptr
    const double *hi_v;
};

struct ciao_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_pa = nullptr, ev_pb = nullptr, ev_sa = nullptr, ev_sb = nullptr, tm_a = nullptr, tm_b = nullptr, ev_pc = nullptr;
    bool pass_timed = false, seq_timed = false, tail_timed = false;
    // problem
    int loss_kind = -1;
    int64_t N_total = 0, row0 = 0, n_rows = 0, d = 0, d_pad = 0, ld = 0;   // n_rows, N_total count COMPONENTS
    // Interleaved row shards (ciao_set_row_interleave): il_block > 0 ⇒ this context holds the blocks of il_block consecutive rows
    // number il_rank, il_rank + il_world, … of the global row sequence, in that order (local row r is global row il_global(r)).
    // Every static minibatch of a multiple of il_block·il_world rows is then spread evenly over the ranks.  Contiguous shards
    // (row0, n_rows) otherwise.
    int64_t il_block = 0;
    int il_rank = 0, il_world = 1;
    int M = 1;                         // rows per component (ciao_set_row_blocks); the record array holds n_rows·M rows
    bool force_block = false;          // env CIAO_FORCE_BLOCK_KERNEL=1: M = 1 problems through the general block kernel too (tests)
    int64_t win0 = 0, win_n = 0;       // pass window over the local rows (0 = all)
    bool win_uniform = false;          // every rank's window is rows [rank·N/world, (rank+1)·N/world): the step scalars can be all-gathered
    double *rec = nullptr;             // row records
    double *qd = nullptr, *ql = nullptr;  // sharing blocks: diag(Q_i), linear term  [N][d_pad]
    double box_lo = 0, box_hi = 0, eta = 0;
    RegParams reg{CIAO_REG_ZERO, 0, 0, 0, nullptr, nullptr};
    double *reg_bounds = nullptr;
    // state
    double *vecs = nullptr;
    double *table = nullptr;
    double *gamma_dev = nullptr;       // [N]
    double *gpair = nullptr;           // ProShI: [N][2] (γ_i, γ_i/N), what one block step needs in 16 bytes
    // adaptive Finito (Finito_adaptive.jl): per-component {γ_i, f_i(x_i), c_i(x_i), 0}, γ̂ on the device, counters of the last call
    double *adapt = nullptr, *adapt_scal = nullptr;
    int64_t *adapt_counters = nullptr;
    double adapt_alpha = 0, adapt_tol_b = 0;
    int64_t adapt_backtracks = 0;
    int algo = 0;                      // 1 svrg, 2 saga, 3 finito, 4 lfinito, 5 proshi
    double gamma = 0, hat_gamma = 0;
    int plus = 0, sag = 0;
    bool cz_valid = false;             // ss holds c_i(z_full) for the current z_full
    bool cz_local_valid = false;       // … at least for the rows of this context (what the LFinito minibatch kernel needs); ss_row0:
    int64_t ss_row0 = 0;               // index in ss of this context's first row (0, or row0 when ss covers all shards' rows)
    // The full-gradient pass at z_full leaves the scalars a sequential step needs, {b_i, λ_i, 0, c_i(z_full)}, in the dense
    // array ss[n_rows][4] (one full 32-byte sector per row, 0.1 % of the pass traffic), so that the SVRG/LFinito step needs
    // one dot product instead of two, forms ∇f_i(z_full) = c_i·a_i while the cluster exchange is in flight, and its producer
    // stages the scalars with one bulk copy.  (A first version wrote c_i into the record tails: 4M scattered 8-byte writes
    // → partial-sector RMW, +1.4 ms per pass.)  Single-process, un-windowed passes only; env CIAO_CACHE_CZ=0 disables it.
    bool cache_cz = true;
    double *ss = nullptr;
    int64_t ss_cap = 0;                // rows the ss allocation holds
    // Debug/test knob (env CIAO_SEQ_TABLE_LDG=1): SAGA/Finito table rows by register prefetch instead of the TMA-staged
    // ring (the path taken anyway when the ring does not fit in shared memory).
    bool seq_table_ldg = false;
    // minibatch sequences run in one persistent cooperative kernel (batch.cu); env CIAO_BATCH_PER_LAUNCH=1 keeps one pass per batch
    bool batch_persistent = true;
    // workspace
    double *ws = nullptr;  size_t ws_bytes = 0;
    double *partial = nullptr;         // [d_pad + 8] partial d-vector + scalars (allreduce buffer)
    int64_t *idx_raw = nullptr, *idx_prep = nullptr, *ptr_dev = nullptr;
    size_t idx_cap = 0, ptr_cap = 0;
    int64_t staged = 0;
    int *err_dev = nullptr;
    int *seq_smid = nullptr;           // [16] SM ids of the CTAs of the last sequential cluster kernel (ciao_last_seq_placement)
    int seq_smid_n = 0;
    int seq_cluster_pos = 0;           // which cluster of a full grid runs the sequential kernels (env CIAO_SEQ_CLUSTER_POS, calibration)
    unsigned int *grid_bar = nullptr;  // grid barrier counter of the persistent minibatch kernel (batch.cu, CIAO_BATCH_EXCHANGE=barrier)
    // flagged-word exchange buffers of the persistent minibatch kernel (batch.cu batch_ll_kernel): [z | Σγ̂/γ_i ×2 | CTA partials], every
    // double stored as two 64-bit words {32 data bits, 32-bit epoch}; zeroed at allocation, the epoch keeps counting across launches
    unsigned long long *ll_buf = nullptr;  size_t ll_bytes = 0;
    int ll_grid = 0;  int64_t ll_d = 0;  uint32_t ll_epoch = 0;
    double *host_pin = nullptr; size_t host_pin_bytes = 0;
    // comm
    void *nccl_comm = nullptr; int rank = 0, world = 1;
    PeerTable peers{};                 // filled by ciao_attach_peer_rows (n = 0: not attached)
    void *peer_mapped[CIAO_MAX_PEERS] = {};  // cudaIpcOpenMemHandle results to close
    // one-shot peer-memory exchange (comm.cu, pass.cu pass_tail_kernel): every rank owns an arena of mail slots + flags that the
    // peers map (CUDA IPC, or directly inside one process) and store into; deterministic rank-ordered sums, no NCCL on the pass path
    double *p2p_arena = nullptr;               // mine
    double *p2p_peer[CIAO_MAX_PEERS] = {};     // every rank's arena as addressed from this GPU (own entry = p2p_arena)
    void *p2p_mapped[CIAO_MAX_PEERS] = {};     // cudaIpcOpenMemHandle results to close
    bool p2p_ready = false;
    uint32_t p2p_seq = 0;                      // collectives issued so far (all ranks in lock step)
    uint32_t p2p_ll_epoch = 0;                 // batches exchanged so far by the persistent minibatch kernel (all ranks in lock step)
    uint64_t p2p_timeout_ns = 20000000000ull;  // a peer that never shows up ends the wait with CIAO_ERR_COMM, not a hang
    // tuning
    int pass_threads = 0, pass_stages = 0, pass_ctas = 0, seq_cluster = 0, seq_threads = 0;
    ciao_timing timing{0, 0, 0, 0, 0, 0};
};

void ciao_set_error(const char *fmt, ...);
#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ciao_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? CIAO_ERR_OOM : CIAO_ERR_CUDA;          \
        }                                                                                     \
    } while (0)
#define CIAO_TRY(expr)            \
    do {                          \
        int _r = (expr);          \
        if (_r != CIAO_OK) return _r; \
    } while (0)
#define CIAO_FAIL(code, ...)        \
    do {                            \
        ciao_set_error(__VA_ARGS__); \
        return (code);              \
    } while (0)

static inline double *ctx_vec(ciao_ctx *c, int which) { return c->vecs + (size_t)which * c->d_pad; }

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// variants taking shared-window addresses computed once (the producer loop of seq_impl.cuh)
__device__ __forceinline__ void mbar_arrive_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra W_%=;\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d_s(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
// 16-byte asynchronous copy global → shared through the GENERIC proxy (LDGSTS, L2 only); used for data that is also written
// inside the same kernel, where a TMA bulk copy (async proxy) would need a proxy fence in the writing thread
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
// the executing thread's earlier cp.async copies arrive on the mbarrier when they complete (the count was armed for it: .noinc)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most n of the executing thread's committed cp.async groups are pending (n ≤ 3 here; the operand is an immediate)
__device__ __forceinline__ void cp_async_wait_pending(int n) {
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void sts_b64(uint32_t addr, int64_t v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
// non-blocking phase test: issued early, its result consumed after independent work (latency ≈ 150 cycles)
__device__ __forceinline__ uint32_t mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra W_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared::cta, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// same, with an L2 evict-first policy for data that is streamed exactly once
__device__ __forceinline__ void tma_load_1d_stream(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar,
                                                   uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2(const void *gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
    return r;
}
// remote shared-memory store whose completion is signalled on a (remote) mbarrier by tx-count:
// no release/acquire fence at cluster scope is needed on either side (SASS: STAS)
__device__ __forceinline__ void st_async_v2f64(uint32_t remote_addr, double a, double b, uint32_t remote_bar_addr) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr),
                 "d"(a), "d"(b), "r"(remote_bar_addr)
                 : "memory");
}
// x / den for a loop-invariant den with rden = 1/den precomputed (correctly rounded): one Newton
// correction of q0 = x·rden with the exact remainder.  Gives the correctly rounded quotient (Markstein)
// without the ~40-instruction division subroutine on the step's critical path.
__device__ __forceinline__ double div_by(double x, double den, double rden) {
    const double q0 = __dmul_rn(x, rden);
    const double rem = fma(-q0, den, x);
    return fma(rem, rden, q0);
}

// Warp-wide sum on the fp64 tensor core: two m8n8k4 DMMAs and one add, no shuffle (measured on B200,
// scripts/micro/dmma_micro.cu: DMMA ≈ 28 cycles, a double shuffle+DADD round ≈ 45 cycles; five shuffle rounds ≈ 225
// cycles, DMMA + SHFL + DMMA + DADD ≈ 90, this version ≈ 65).  Used where the reduction sits on a sequential critical path.
// Fragment layout of mma.m8n8k4.f64: lane L holds A[L>>2][L&3], B[L&3][L>>2] and D[L>>2][2(L&3) + {0,1}].
//   1st DMMA:  A = 1, B[k][j] = v of lane 4j+k  →  D[·][j] = S_j = Σ_k v_{4j+k}: lane L holds S_{2c}, S_{2c+1}, c = L&3
//   add:       x = S_{2c} + S_{2c+1}, so the four lanes of every group hold the four pair sums
//   2nd DMMA:  A[i][k] = x of lane 4i+k = S_{2k} + S_{2k+1}, B = 1  →  D[·][·] = Σ_k (S_{2k} + S_{2k+1}) = total
// The result is identical in all 32 lanes and, for identical inputs, in all warps (fixed hardware summation order).
__device__ __forceinline__ void dmma_884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}
__device__ __forceinline__ double warp_sum_mma(double v, int lane) {
    (void)lane;
    double s0, s1, t0, t1;
    dmma_884(s0, s1, 1.0, v);
    dmma_884(t0, t1, s0 + s1, 1.0);
    return t0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// operator device functions (ProximalOperators.jl 0.14 semantics)
// ---------------------------------------------------------------------------
// prox!(y, g, x, γ) for one coordinate.  NormL1: gl = γλ; y = x + (x ≤ −gl ? gl : (x ≥ gl ? −gl : −x))
template <int REG>
__device__ __forceinline__ double prox_elem(double x, double gl, double lo, double hi) {
    if (REG == CIAO_REG_NORML1) {
        // sign(x)·max(|x| − gl, 0) assembled with integer selects: the same bits as the reference's
        // x + (x ≤ −gl ? gl : (x ≥ gl ? −gl : −x)) — |x| − gl rounds like x − gl and, negated, like x + gl; the result is
        // +0 whenever |x| ≤ gl (strict test of the 64-bit pattern), NaN propagates — with ONE instruction on the fp64 pipe and
        // a 3-deep dependency chain instead of DSETP/FSEL/DSETP/FSEL/DADD (the step is fp64-issue bound after the exchange).
        // (x − clamp(x, −gl, gl) with DMNMX is the short form, but DMNMX is slow on B200: 0.40 vs 0.377 µs/step.)
        const double t = __dsub_rn(fabs(x), gl);
        const long long tb = __double_as_longlong(t);
        const int sx = __double2hiint(x) & 0x80000000;
        const bool pos = tb > 0;   // t > +0 (or NaN with the sign bit clear)
        return __hiloint2double(pos ? (__double2hiint(t) | sx) : 0, pos ? __double2loint(t) : 0);
    } else if (REG == CIAO_REG_INDBOX) {
        return x < lo ? lo : (x > hi ? hi : x);
    }
    return x;
}

// prox_g with the kind chosen at run time (d-sized helper kernels; the step kernels template on REG)
__device__ __forceinline__ double prox_rt(int kind, double x, double gl, double lo, double hi) {
    if (kind == CIAO_REG_NORML1) return prox_elem<CIAO_REG_NORML1>(x, gl, lo, hi);
    if (kind == CIAO_REG_INDBOX) return prox_elem<CIAO_REG_INDBOX>(x, gl, lo, hi);
    return x;
}

// scalar c(u) with ∇f_i(x) = c · a_i  and the value f_i(x), from u = a_i·x
//   LS:       res = u − b;  ∇ = (a·res)·λ  (two roundings per element, as gradient! does);  f = (λ/2)·res²
//   logistic: c = −μ y /(1 + exp(y u));  ∇ = a·c;  f = μ·log(1 + 1/exp(y u))
template <int LOSS>
__device__ __forceinline__ double loss_coef(double u, double b, double lam) {
    if (LOSS == CIAO_LOSS_LS) return __dsub_rn(u, b);  // residual; λ applied per element
#ifdef CIAO_LIBM_LOGISTIC
    double e = exp(__dmul_rn(b, u));  // CUDA math library: ≈ 430 cycles of dependent latency per step
    return __ddiv_rn(__dmul_rn(-lam, b), __dadd_rn(1.0, e));
#else
    return logistic_coef_fast(u, b, lam);  // fastmath.cuh: same formula, latency-optimised exp and division
#endif
}
// c_p = loss_coef(u_p, b_p, λ_p) for p = 0 … P−1 with ONE evaluation per warp: lane p takes pair p, the results come back by
// shuffle.  The logistic coefficient is ≈ 60 instructions; with every warp evaluating all P (row, dot) pairs of a row group the
// streaming passes at d = 1024 were instruction-bound in exactly that code (LFinito minibatch pass: 3.3 TB/s).  Same function on
// the same inputs: bit-identical to P separate evaluations.  Must be called by all 32 lanes.
template <int LOSS, int P>
__device__ __forceinline__ void loss_coef_lanes(const double (&u)[P], const double (&b)[P], const double (&lam)[P], int lane,
                                                double (&c)[P]) {
    static_assert(P <= 32, "one lane per pair");
    if (LOSS == CIAO_LOSS_LS || P == 1) {
#pragma unroll
        for (int q = 0; q < P; ++q) c[q] = loss_coef<LOSS>(u[q], b[q], lam[q]);
        return;
    }
    double ul = u[0], bl = b[0], ll = lam[0];
#pragma unroll
    for (int q = 1; q < P; ++q) {
        const bool m = lane == q;
        ul = m ? u[q] : ul;
        bl = m ? b[q] : bl;
        ll = m ? lam[q] : ll;
    }
    const double cl = loss_coef<LOSS>(ul, bl, ll);
#pragma unroll
    for (int q = 0; q < P; ++q) c[q] = __shfl_sync(0xffffffffu, cl, q);
}
template <int LOSS>
__device__ __forceinline__ double loss_value(double u, double b, double lam) {
    if (LOSS == CIAO_LOSS_LS) {
        double r = u - b;
        return (lam / 2) * (r * r);
    }
    double e = exp(b * u);
    return lam * log(1 + 1 / e);
}
// one gradient coordinate in the reference's rounding order
template <int LOSS>
__device__ __forceinline__ double grad_elem(double a, double c, double lam) {
    if (LOSS == CIAO_LOSS_LS) return __dmul_rn(__dmul_rn(a, c), lam);
    return __dmul_rn(a, c);
}
#endif  // __CUDACC__

/* ciao_cuda.h — C ABI of libciao_cuda, the B200 (sm_100a) engine for the
 * iteration loops of kul-optec/CIAOAlgorithms.jl.
 *
 * The reference has no FFI today: its hot path sits behind two Julia generic
 * protocols — the iteration protocol Base.iterate(iter) /
 * Base.iterate(iter, state) / solution(state) and the ProximalOperators
 * gradient!/prox! protocol.  Each entry point below replaces the body of one
 * of those methods for the operator kinds the reference's tests use; the
 * Julia-side `ccall` stubs are shown in INTEGRATION.md and julia/CIAOAlgorithmsCUDA.jl.
 *
 * Conventions
 *   - every function returns 0 on success or a negative CIAO_ERR_* code; the
 *     message is available from ciao_last_error() (thread-local).  Nothing
 *     throws or aborts across the boundary (reference behaviour: @warn +
 *     `return nothing`, e.g. SVRG_basic.jl:36-42).
 *   - host pointers are borrowed for the duration of the call only.
 *   - indices cross the boundary exactly as Julia produces them: int64, 1-based;
 *     they are range-checked (CIAO_ERR_INVALID), never trusted.
 *   - `idx` arguments may be host pointers, device pointers, or NULL (= use the
 *     indices last staged with ciao_stage_indices).
 *   - one ciao_ctx owns one GPU and is not thread-safe; contexts are independent.
 *   - all arithmetic is IEEE fp64.
 *   - there is no CPU fallback: without a GPU, ciao_create fails with CIAO_ERR_CUDA.
 */
#ifndef CIAO_CUDA_H
#define CIAO_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CIAO_VERSION 100 /* 0.1.0 */

#define CIAO_OK 0
#define CIAO_ERR_INVALID (-1)     /* bad argument (shape, kind, index out of range)   */
#define CIAO_ERR_CUDA (-2)        /* CUDA runtime/driver error, or no GPU             */
#define CIAO_ERR_STATE (-3)       /* call order: problem/solver state not initialised */
#define CIAO_ERR_UNSUPPORTED (-4) /* operator kind / size outside the engine's scope  */
#define CIAO_ERR_COMM (-5)        /* NCCL / peer-memory error                         */
#define CIAO_ERR_OOM (-6)         /* device allocation failed                         */

/* f_i kinds (ProximalOperators objects recognised by the shim) */
#define CIAO_LOSS_LS 0       /* LeastSquares(a_i' (1×d), [b_i], λ_i):  (λ_i/2)(a_i·x − b_i)²      test_lasso.jl:53-54 */
#define CIAO_LOSS_LOGISTIC 1 /* Precompose(LogisticLoss([y_i], μ_i), a_i', 1): μ_i log(1+exp(−y_i a_i·x))  test_logistic_l1.jl:36 */
#define CIAO_LOSS_DIAGQUAD 2 /* Sum(Quadratic(diag(q_i), c_i), SqrDistL2(IndBox(lo,hi), η))      test_sharing.jl:18-22 */

/* g kinds */
#define CIAO_REG_ZERO 0   /* Zero()            prox = identity            SVRG.jl:49   */
#define CIAO_REG_NORML1 1 /* NormL1(λ)         soft threshold             test_lasso.jl:59 */
#define CIAO_REG_INDBOX 2 /* IndBox(lo, hi)    clamp (scalar or vector)   test_sharing.jl:25 */
#define CIAO_REG_NORML1_PAIRS 3 /* NormL1(λ) on complex data stored as interleaved (re, im) pairs: sign(x)·max(0, |x| − γλ)
                                   (ProximalOperators normL1.jl, complex method; test_lasso.jl:3 with T = ComplexF64)          */

/* state vectors readable with ciao_get_vec (names follow the reference's state structs) */
#define CIAO_VEC_Z 0      /* state.z      (SAGA/Finito/LFinito/ProShI iterate; SVRG inner sum) */
#define CIAO_VEC_Z_FULL 1 /* state.z_full (SVRG snapshot, LFinito)                             */
#define CIAO_VEC_W 2      /* state.w      (SVRG inner iterate)                                 */
#define CIAO_VEC_AV 3     /* state.av     (running average / full gradient)                    */
#define CIAO_VEC_X 4      /* scratch: last x handed to a pass                                  */

/* synthetic problems (include/ciao_gen.h) */
#define CIAO_SYNTH_LASSO 0
#define CIAO_SYNTH_LOGISTIC 1
#define CIAO_SYNTH_SHARING 2

typedef struct ciao_ctx ciao_ctx;

typedef struct {
    float last_pass_ms;     /* device time of the last streaming pass kernel (CUDA events on the ctx stream) */
    float last_seq_ms;      /* device time of the last sequential (persistent) kernel                       */
    int64_t last_pass_bytes; /* algorithmic HBM bytes of that pass                                           */
    int64_t last_seq_steps;
    int64_t launches;       /* kernels launched by this context so far                                      */
    float last_tail_ms;     /* device time of the last pass's tail kernel: CTA-partial reduction, exchange over the
                               ranks (including the wait for the slowest rank) and the closing update              */
} ciao_timing;

/* ---- lifetime ----------------------------------------------------------- */
int ciao_version(void);
const char *ciao_last_error(void);
int ciao_device_count(int *n);
int ciao_create(ciao_ctx **out, int device);
int ciao_destroy(ciao_ctx *ctx);
int ciao_sync(ciao_ctx *ctx);

/* ---- problem: F = [f_1..f_N], g ------------------------------------------ */
/* Row models (replaces the packing of F::Array{Tf}, SVRG_basic.jl:2).  A is
 * row-major n_rows×d with leading dimension lda (== a Julia column-major d×N
 * matrix); b_or_y and scale (λ_i or μ_i; NULL → scale_scalar) have n_rows
 * entries.  The rows are this context's shard [row0, row0+n_rows) of a problem
 * with N_total components (single GPU: row0 = 0, n_rows = N_total). */
int ciao_set_rows(ciao_ctx *ctx, int loss_kind, int64_t N_total, int64_t row0, int64_t n_rows, int64_t d,
                  const double *A, int64_t lda, const double *b_or_y, const double *scale, double scale_scalar);
/* Components that are M×d blocks: f_i = LeastSquares(A_i (M×d), b_i (M), λ_i) or Precompose(LogisticLoss(y_i (M), μ_i), L_i (M×d))
 * (test_lasso.jl:52-54 is the case M = 1).  A is row-major (N·M)×d, component i holds rows i·M … i·M+M−1; b_or_y has N·M
 * entries, scale N entries (NULL → scale_scalar).  A complex 1×d row is the M = 2 real block [Re; Im] of the realified problem
 * (x interleaved as re, im), to be used with CIAO_REG_NORML1_PAIRS.  The passes stream the N·M rows like any rows; the sequential
 * loops run in the general block kernel (csrc/blockseq.cu), not in the tuned cluster kernels; not sharded, no adaptive Finito. */
int ciao_set_row_blocks(ciao_ctx *ctx, int loss_kind, int64_t N, int64_t M, int64_t d, const double *A, int64_t lda,
                        const double *b_or_y, const double *scale, double scale_scalar);
/* Interleaved row shards (one process per GPU; the reference has no counterpart — where the rows of F live is the host's
 * choice): rank k of `world` holds the blocks of block_rows consecutive rows number k, k + world, … in that order.  Call it
 * before the rows are set; ciao_set_rows / ciao_gen_synthetic then take row0 = rank·block_rows and n_rows = the number of rows
 * the rank owns.  A static minibatch (Finito_basic.jl:52-57: contiguous rows) that starts at a multiple of block_rows·world
 * is then spread evenly over the ranks, so minibatch sweeps scale with the GPUs; passes shard as with contiguous shards.
 * Single-sample steps and remote rows are for contiguous shards.  block_rows = 0 returns to contiguous shards. */
int ciao_set_row_interleave(ciao_ctx *ctx, int64_t block_rows, int rank, int world);
/* Sharing blocks f_i(x_i) = ½x'diag(q_i)x + c_i'x + (η/2)dist²(x, [lo,hi])   (test_sharing.jl:15-22) */
int ciao_set_blocks(ciao_ctx *ctx, int64_t N, int64_t n, const double *Qdiag, int64_t ldq, const double *qlin,
                    int64_t ldl, double box_lo, double box_hi, double eta);
/* g: ZERO (nparams 0) | NORML1 (params = {λ}) | INDBOX (params = {lo,hi} or lo[d] followed by hi[d]) | NORML1_PAIRS (params = {λ}) */
int ciao_set_reg(ciao_ctx *ctx, int reg_kind, const double *params, int64_t nparams);
/* Counter-based synthetic shard generated directly in HBM (ciao_gen.h); scale = λ_i / μ_i for all rows */
int ciao_gen_synthetic(ciao_ctx *ctx, int synth_kind, int64_t N_total, int64_t row0, int64_t n_rows, int64_t d,
                       uint64_t seed, double scale);
/* Host twin of the generator (same bits), for callers that need the rows on the host */
int ciao_gen_host(int synth_kind, int64_t d, uint64_t seed, int64_t row0, int64_t n_rows, double *A, double *rhs);

/* ---- multi-GPU: one process per GPU, row-sharded passes -------------------- */
/* out must hold 128 bytes (ncclUniqueId); rank 0 creates it, the host broadcasts it */
int ciao_comm_unique_id(void *out128);
int ciao_comm_init(ciao_ctx *ctx, const void *id128, int rank, int world);
/* One-shot exchange over peer memory — the deterministic all-reduce of the row-sharded passes (SVRG_basic.jl:58-63 over a
 * sharded F).  Every rank exports an arena (ciao_comm_p2p_handle: 128 opaque bytes), the host all-gathers
 * the blobs in rank order and every rank attaches them (CUDA IPC between processes — also between processes sharing one GPU —
 * or plain peer access inside one process).  From then on the tail kernel of a pass stores its partial d-vector into every
 * peer's arena, raises a flag, waits for the peers' flags (bounded: a missing peer gives CIAO_ERR_COMM at the next synchronising
 * call) and sums the slots in rank order: bit-identical results on every rank and run to run, one kernel launch after the pass.
 * Without attached arenas the passes fall back to ncclAllReduce (ciao_comm_init).  Collective calls must be made by all
 * ranks in the same order. */
int ciao_comm_p2p_handle(ciao_ctx *ctx, void *out128);
int ciao_comm_p2p_attach(ciao_ctx *ctx, int rank, int world, const void *handles128);

/* Replicated data, sharded pass: restrict the full-gradient / objective passes of this context to
 * the local rows [row_lo, row_lo + n) (0-based); with a communicator the partial d-vectors are
 * all-reduced, so G ranks holding the same rows each stream 1/G of them.  n = 0 resets.  With a communicator the call is
 * COLLECTIVE (all ranks must make it): the ranks agree on whether the windows tile [0, N) uniformly, in which case a
 * full-gradient pass also all-gathers the per-row scalars c_i(z_full) its inner epochs use. */
int ciao_set_pass_window(ciao_ctx *ctx, int64_t row_lo, int64_t n);

/* Row-sharded problems beyond one GPU's HBM: every rank exports its shard (ciao_rows_ipc_handle, 64 bytes, CUDA IPC),
 * the host all-gathers the handles and shard bounds, and each rank attaches the peers' shards.  The sequential inner
 * epochs without a table (SVRG/SVRG++, LFinito) then run replicated on every rank, TMA-prefetching remote rows over
 * NVLink; the passes stay sharded + all-reduced.  Shards must tile [0, N) in rank order. */
int ciao_rows_ipc_handle(ciao_ctx *ctx, void *out64);
int ciao_attach_peer_rows(ciao_ctx *ctx, int n_shards, const void *handles64, const int64_t *row0, const int64_t *n_rows,
                          int my_shard);

/* ---- streaming passes (HBM-bound) ------------------------------------------ */
/* out = scale · Σ_i ∇f_i(x)      (SVRG_basic.jl:58-63, 88-92; Finito_LFinito.jl:68-72, 85-88) */
int ciao_full_gradient(ciao_ctx *ctx, const double *x, double scale, double *out_or_null);
/* f_mean = (1/N) Σ f_i(x), g_val = g(x)      (cost of test_lasso.jl:45) */
int ciao_objective(ciao_ctx *ctx, const double *x, double *f_mean, double *g_val);
/* max_i ‖a_i‖² — for L_i = λ_i‖a_i‖² (test_lasso.jl:55) / 0.25‖a_i‖² (test_logistic_l1.jl:39) */
int ciao_max_row_sqnorm(ciao_ctx *ctx, double *out);

/* ---- SVRG / SVRG++  (SVRG_basic.jl) ---------------------------------------- */
int ciao_svrg_init(ciao_ctx *ctx, const double *x0, double gamma, int plus);   /* :58-66 */
int ciao_svrg_epoch(ciao_ctx *ctx, const int64_t *idx, int64_t m);             /* :73-92 (caller doubles m, :93) */

/* ---- SAGA / SAG  (SAGA_basic.jl) ------------------------------------------- */
int ciao_saga_init(ciao_ctx *ctx, const double *x0, double gamma, int sag);    /* :41-48 */
int ciao_saga_steps(ciao_ctx *ctx, const int64_t *idx, int64_t K);             /* :53-68, K times */

/* ---- Finito / MISO / DIAG and LFinito  (Finito_basic.jl, Finito_LFinito.jl) -- */
/* hat_gamma is computed by the caller in the reference's order (1/sum(1 ./ γ)) */
int ciao_finito_init(ciao_ctx *ctx, const double *x0, const double *gamma_N, double hat_gamma);  /* Finito_basic.jl:76-84 */
/* batch j = idx[batch_ptr[j] .. batch_ptr[j+1]) ; the prox closes every batch (:110-118) */
int ciao_finito_steps(ciao_ctx *ctx, const int64_t *idx, const int64_t *batch_ptr, int64_t n_batches);
int ciao_lfinito_init(ciao_ctx *ctx, const double *x0, const double *gamma_N, double hat_gamma); /* Finito_LFinito.jl:67-72 */
/* one outer iteration (:78-103): batch_order = state.inds (1-based), static batches of r rows */
int ciao_lfinito_outer(ciao_ctx *ctx, const int64_t *batch_order, int64_t n_batches, int64_t r);

/* ---- Finito adaptive  (Finito_adaptive.jl) ----------------------------------- */
/* :59-99 — tables x_i = x0, ∇f_i(x0) (kept as the scalar c_i: ∇f_i = c_i·a_i for row models), f_i(x0);
 * γ_i = α / (‖∇f_i(x0+1) − ∇f_i(x0)‖ / (√d·N)); γ̂, av, z.  CIAO_ERR_UNSUPPORTED if some ∇f_i(x0+1) == ∇f_i(x0)
 * (the reference then draws random perturbations from the global RNG, :75-81: use ciao_finito_adaptive_init_cb). */
int ciao_finito_adaptive_init(ciao_ctx *ctx, const double *x0, double alpha, double tol_b);
/* Same, with the random restart of :77-83: for a component i with ∇f_i(x0+1) == ∇f_i(x0) the library calls
 * perturb(user, i (1-based), t, xeps), which must fill xeps[0..d) = x0 .+ rand(t·[−1, 1], size(x0)) from the HOST's RNG (Julia:
 * a @cfunction) and return 0 — components in ascending order, t = 1, 2, 4, … per component: the reference's draw order. */
typedef int (*ciao_perturb_fn)(void *user, int64_t i, int64_t t, double *xeps);
int ciao_finito_adaptive_init_cb(ciao_ctx *ctx, const double *x0, double alpha, double tol_b, ciao_perturb_fn perturb, void *user);
/* :101-160, K single-index steps with the backtracking linesearch on γ_i; *steps_done < K ⇔ the reference's
 * `return nothing` (γ_i < tol_b/N, :124-127) at step *steps_done + 1 */
int ciao_finito_adaptive_steps(ciao_ctx *ctx, const int64_t *idx, int64_t K, int64_t *steps_done);
/* state.γ, state.fi_x, c_i (N entries each, any may be NULL), state.hat_γ, number of 0.8-reductions so far */
int ciao_finito_adaptive_get(ciao_ctx *ctx, double *gamma_N, double *fi_x_N, double *coef_N, double *hat_gamma,
                             int64_t *backtracks);

/* ---- ProShI  (ProShI_basic.jl) --------------------------------------------- */
int ciao_proshi_init(ciao_ctx *ctx, const double *x0, const double *gamma_N, double hat_gamma);  /* :76-86 */
int ciao_proshi_steps(ciao_ctx *ctx, const int64_t *idx, const int64_t *batch_ptr, int64_t n_batches); /* :111-123 */
/* applies s_i += γ_i z IN PLACE on every call, like the reference (:127-132); S_out N×n row-major or NULL */
int ciao_proshi_solution(ciao_ctx *ctx, double *S_out_or_null);

/* ---- state access ("the iterator state is the checkpoint") ------------------ */
int ciao_get_vec(ciao_ctx *ctx, int which, double *out, int64_t len);
int ciao_set_vec(ciao_ctx *ctx, int which, const double *in, int64_t len);
int ciao_get_table_rows(ciao_ctx *ctx, int64_t i0, int64_t n, double *out);   /* rows [i0, i0+n) of s, 0-based */
int ciao_set_table_rows(ciao_ctx *ctx, int64_t i0, int64_t n, const double *in); /* restore rows [i0, i0+n) of s        */
/* Restore from a checkpoint: re-creates a solver's device state without its init pass (SAGA_basic.jl:11-20 — the state
 * struct IS the checkpoint): algo 1 SVRG (gamma, flag = plus) | 2 SAGA (gamma, flag = SAG) | 3 Finito | 4 LFinito | 5 ProShI
 * (gamma_N, hat_gamma); then ciao_set_vec (z, z_full, w, av) and ciao_set_table_rows put the state back. */
int ciao_solver_restore(ciao_ctx *ctx, int algo, double gamma, int flag, const double *gamma_N, double hat_gamma);
int ciao_table_colsum(ciao_ctx *ctx, double *out);                             /* Σ_i s_i (sum(x_proshi), test_sharing.jl:42) */

/* ---- host-side index draws (no device involved) ----------------------------- */
/* Restatement of Julia >= 1.7's default RNG (task-local Xoshiro256++) and of the samplers the reference's RNG call sites go
 * through (SVRG_basic.jl:73 rand(ind, m); SAGA_basic.jl:55 rand(1:N); Finito_basic.jl:97 sample(1:N, k, replace=false);
 * Finito_basic.jl:102 randperm(d)), for hosts without Julia (the Python twin of the shim).  state4 = the four UInt64 state
 * words (in/out); seeding (Random.seed!(n): SHA-256 of the seed's UInt32 words) is the caller's.  The raw stream is pinned by
 * the known answer in Julia's documentation; the integer samplers are restated from Random / StatsBase 0.33 sources and are
 * unverified against a running Julia (jlrng.cu).  All outputs are 1-based, as Julia returns them. */
int ciao_jlrng_next_u64(uint64_t *state4, uint64_t *out, int64_t n);
int ciao_jlrng_rand_range(uint64_t *state4, int64_t N, int64_t *out, int64_t m);
int ciao_jlrng_randperm(uint64_t *state4, int64_t n, int64_t *out);
int ciao_jlrng_sample_norep(uint64_t *state4, int64_t N, int64_t k, int64_t *out);

/* ---- measurement ----------------------------------------------------------- */
int ciao_stage_indices(ciao_ctx *ctx, const int64_t *idx_host, int64_t n);    /* pre-upload indices (device-resident timing) */
int ciao_timer_begin(ciao_ctx *ctx);                                          /* CUDA event on the ctx stream */
int ciao_timer_end(ciao_ctx *ctx, float *ms);                                 /* records, synchronises, returns elapsed */
int ciao_last_timing(ciao_ctx *ctx, ciao_timing *out);
/* SM ids of the CTAs of the last sequential cluster kernel (n_ctas ≤ 16 entries of smid16 are valid) */
int ciao_last_seq_placement(ciao_ctx *ctx, int *smid16, int *n_ctas);
/* SM cycles and nanoseconds the last sequential cluster kernel ran: cycles / ns = the SM clock (GHz) it actually saw */
int ciao_last_seq_clock(ciao_ctx *ctx, int64_t *cycles, int64_t *ns);
/* Latency floor of the sequential kernels' cluster exchange on this device (seq_floor.cu): every cluster of `cluster` CTAs ×
 * `warps` warps that fits on the GPU runs `iters` rounds of the st.async → mbarrier exchange (mode 0) or of exchange + partial
 * sum + dependent next message (mode 1) and reports its own ns and SM cycles per round and the SMs it sits on
 * (smid[k·cluster + r]).  Arrays hold max_clusters (·cluster) entries; *n_clusters of them are written. */
int ciao_measure_exchange(ciao_ctx *ctx, int cluster, int warps, int iters, int mode, int max_clusters, int *n_clusters,
                          float *ns_per_round, float *cyc_per_round, int *smid);
/* tuning knobs (0 = default): threads per CTA and ring stages of the streaming pass, cluster size of the sequential kernels */
int ciao_set_tuning(ciao_ctx *ctx, int pass_threads, int pass_stages, int pass_ctas_per_sm, int seq_cluster, int seq_threads);

#ifdef __cplusplus
}
#endif
#endif /* CIAO_CUDA_H */

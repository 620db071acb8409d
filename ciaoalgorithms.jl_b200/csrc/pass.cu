// pass.cu — K1/K2/K9: the HBM-bound streaming passes over all row records.
//
//   PASS_GRAD        out = Σ_i ∇f_i(x)            SVRG_basic.jl:58-63, 88-92; Finito_LFinito.jl:68-72, 85-88
//   PASS_SAGA_INIT   s_i = ∇f_i(x0); Σ s_i        SAGA_basic.jl:41-47
//   PASS_FINITO_INIT s_i = x0 − (γ_i/N)∇f_i(x0); Σ s_i/γ_i     Finito_basic.jl:76-83
//   PASS_NORMS       max_i ‖a_i‖²                 (L_i of test_lasso.jl:55, test_logistic_l1.jl:39)
// every mode also yields Σ_i f_i(x) (the objective of test_lasso.jl:45) for free.
//
// Design: one persistent CTA per SM streams contiguous groups
// of RPG row records (≈32 KB) through an S-stage shared-memory ring filled by
// 1-D TMA bulk copies (cp.async.bulk → UBLKCP) with an L2 evict-first policy.
// A is read from HBM exactly once: a thread owns CPT fixed columns, pulls them
// from the ring into registers, the row dot a_i·x is reduced warp-shuffle →
// shared → all threads (fixed order), and the axpy  acc += c_i·a_i  reuses the
// registers.  CTA partial d-vectors go to a workspace and are summed by
// reduce_ws_kernel in a fixed order: no floating-point atomics anywhere, so the
// result is bitwise reproducible run to run (test_lasso.jl:192 `==` tests).
#include <string.h>

#include <algorithm>

#include "common.cuh"

enum { PASS_GRAD = 0, PASS_SAGA_INIT = 1, PASS_FINITO_INIT = 2, PASS_NORMS = 3 };

struct PassArgs {
    const double *rec;   // [n_rows][ld]
    double *ss_out;      // dense [n_rows][4] ← {b_i, λ_i, 0, c_i(x)}, one 32-byte sector per row (nullptr: do not cache)
    int64_t n_rows, ld, d_pad;
    const double *x;     // [d_pad]
    double *ws;          // [grid][d_pad]
    double *fws;         // [grid]
    double *table;       // [n_rows][d_pad] or nullptr
    double Nd;           // (double) N_total
    int stages;
};

template <int CPT, int MODE, int LOSS>
__global__ void __launch_bounds__(512, 1) row_pass_kernel(const PassArgs p) {
    constexpr int RPG = 16 / CPT;  // rows per group: RPG·CPT = 16 doubles of row data per thread
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.stages;
    const size_t stage_doubles = (size_t)RPG * p.ld;
    double *ring = reinterpret_cast<double *>(smem_raw);
    double *red = ring + (size_t)S * stage_doubles;  // [2][RPG][32]
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 2 * RPG * 32);

    const int64_t n_groups = (p.n_rows + RPG - 1) / RPG;
    const int64_t my_count = (blockIdx.x < n_groups) ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    uint64_t policy = 0;
    for (int i = tid; i < 2 * RPG * 32; i += T) red[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        policy = l2_policy_evict_first();
    }
    __syncthreads();

    auto issue = [&](int64_t it) {
        const int64_t g = blockIdx.x + it * (int64_t)gridDim.x;
        const int64_t r0 = g * RPG;
        const int rows = (int)min((int64_t)RPG, p.n_rows - r0);
        const uint32_t bytes = (uint32_t)(rows * p.ld * sizeof(double));
        const int slot = (int)(it % S);
        mbar_arrive_expect_tx(&full[slot], bytes);
        tma_load_1d_stream(ring + (size_t)slot * stage_doubles, p.rec + r0 * p.ld, bytes, &full[slot], policy);
    };
    if (tid == 0)
        for (int64_t it = 0; it < my_count && it < S; ++it) issue(it);

    // this thread's columns: double2 units u = tid + T·k
    double xr[CPT], acc[CPT];
    int col[CPT / 2];
#pragma unroll
    for (int k = 0; k < CPT / 2; ++k) {
        col[k] = 2 * (tid + T * k);
        const bool v = col[k] < p.d_pad;
        if (!v) col[k] = -1;
        xr[2 * k] = (v && MODE != PASS_NORMS) ? p.x[col[k]] : 0.0;
        xr[2 * k + 1] = (v && MODE != PASS_NORMS) ? p.x[col[k] + 1] : 0.0;
        acc[2 * k] = acc[2 * k + 1] = 0.0;
    }
    double fsum = 0.0;  // thread 0: Σ f_i (or max ‖a_i‖²)

    for (int64_t it = 0; it < my_count; ++it) {
        const int slot = (int)(it % S);
        const uint32_t parity = (uint32_t)((it / S) & 1);
        const int par = (int)(it & 1);
        const int64_t g = blockIdx.x + it * (int64_t)gridDim.x;
        const int64_t r0 = g * RPG;
        const int rows = (int)min((int64_t)RPG, p.n_rows - r0);
        mbar_wait(&full[slot], parity);
        const double *sp = ring + (size_t)slot * stage_doubles;

        double a[RPG][CPT], pd[RPG], tb[RPG], tl[RPG], tg[RPG], tgn[RPG];
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
            pd[r] = 0.0;
            const bool rv = r < rows;
            const double *rp = sp + (size_t)r * p.ld;
#pragma unroll
            for (int k = 0; k < CPT / 2; ++k) {
                double2 v = make_double2(0.0, 0.0);
                if (rv && col[k] >= 0) v = *reinterpret_cast<const double2 *>(rp + col[k]);
                a[r][2 * k] = v.x;
                a[r][2 * k + 1] = v.y;
                if (MODE == PASS_NORMS) {
                    pd[r] = fma(v.x, v.x, pd[r]);
                    pd[r] = fma(v.y, v.y, pd[r]);
                } else {
                    pd[r] = fma(v.x, xr[2 * k], pd[r]);
                    pd[r] = fma(v.y, xr[2 * k + 1], pd[r]);
                }
            }
            tb[r] = rv ? rp[p.d_pad] : 0.0;
            tl[r] = rv ? rp[p.d_pad + TAIL_LAM] : 0.0;
            tg[r] = (rv && MODE == PASS_FINITO_INIT) ? rp[p.d_pad + TAIL_GAM] : 1.0;
            tgn[r] = (rv && MODE == PASS_FINITO_INIT) ? rp[p.d_pad + TAIL_GAM_N] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
            pd[r] = warp_sum_mma(pd[r], lane);   // tensor-core sum: 3 instructions instead of 5 shuffle+add rounds
            if (lane == 0) red[(par * RPG + r) * 32 + warp] = pd[r];
        }
        __syncthreads();
        // every thread has its slice of the stage in registers: the slot can be refilled
        if (tid == 0 && it + S < my_count) issue(it + S);

        // the W warp partials (entries ≥ W stay zero) summed by the same tensor-core reduction in every warp → identical bits
        double uu[RPG], cc[RPG];
#pragma unroll
        for (int r = 0; r < RPG; ++r) uu[r] = warp_sum_mma(red[(par * RPG + r) * 32 + lane], lane);
        if (MODE != PASS_NORMS) loss_coef_lanes<LOSS, RPG>(uu, tb, tl, lane, cc);   // rows beyond `rows` carry zeros: harmless
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
            const double u = uu[r];
            if (r >= rows) continue;
            if (MODE == PASS_NORMS) {
                if (tid == 0) fsum = fmax(fsum, u);
                continue;
            }
            const double c = cc[r];
            if (tid == 0) {
                fsum += loss_value<LOSS>(u, tb[r], tl[r]);
                if (MODE == PASS_GRAD && p.ss_out) {
                    double2 *o = reinterpret_cast<double2 *>(p.ss_out + 4 * (r0 + r));
                    o[0] = make_double2(tb[r], tl[r]);
                    o[1] = make_double2(0.0, c);
                }
            }
            if (MODE == PASS_GRAD) {
                const double cc = (LOSS == CIAO_LOSS_LS) ? c * tl[r] : c;
#pragma unroll
                for (int e = 0; e < CPT; ++e) acc[e] = fma(cc, a[r][e], acc[e]);
            } else {
                const double cg = tgn[r];  // γ_i / N, precomputed in the record tail
                const double rg = (MODE == PASS_FINITO_INIT) ? __ddiv_rn(1.0, tg[r]) : 0.0;  // one division per row, not per element
                double *trow = p.table + (r0 + r) * p.d_pad;
#pragma unroll
                for (int k = 0; k < CPT / 2; ++k) {
                    double s0 = grad_elem<LOSS>(a[r][2 * k], c, tl[r]);
                    double s1 = grad_elem<LOSS>(a[r][2 * k + 1], c, tl[r]);
                    if (MODE == PASS_FINITO_INIT) {
                        s0 = __dsub_rn(xr[2 * k], __dmul_rn(cg, s0));
                        s1 = __dsub_rn(xr[2 * k + 1], __dmul_rn(cg, s1));
                        acc[2 * k] += div_by(s0, tg[r], rg);
                        acc[2 * k + 1] += div_by(s1, tg[r], rg);
                    } else {
                        acc[2 * k] += s0;
                        acc[2 * k + 1] += s1;
                    }
                    if (col[k] >= 0) __stcs(reinterpret_cast<double2 *>(trow + col[k]), make_double2(s0, s1));
                }
            }
        }
    }

    if (MODE != PASS_NORMS) {
        double *wrow = p.ws + (size_t)blockIdx.x * p.d_pad;
#pragma unroll
        for (int k = 0; k < CPT / 2; ++k)
            if (col[k] >= 0) *reinterpret_cast<double2 *>(wrow + col[k]) = make_double2(acc[2 * k], acc[2 * k + 1]);
    }
    if (tid == 0) p.fws[blockIdx.x] = fsum;
}

// second stage: fixed-order sum of the CTA partials.  out[j] = Σ_b ws[b][j];  fout = Σ_b fws[b] (or max).
// Block = 32 columns × REDUCE_SLICES slices of b: thread (x, y) sums b = y, y+S, … (coalesced 256-byte rows),
// the slices are then combined in a fixed order through shared memory — deterministic, and a few µs instead of
// a G-long serial chain of global loads per column (which dominated the minibatch passes).
constexpr int REDUCE_SLICES = 16;
__global__ void __launch_bounds__(32 * REDUCE_SLICES) reduce_ws_kernel(const double *ws, const double *fws, int G, int64_t d_pad,
                                                                       double *out, double *fout, int fmax_mode, int with_vec) {
    __shared__ double sm[REDUCE_SLICES][33];
    const int x = threadIdx.x, y = threadIdx.y;
    const int64_t j = blockIdx.x * 32ll + x;
    if (with_vec) {
        double s = 0.0;
        if (j < d_pad)
            for (int b = y; b < G; b += REDUCE_SLICES) s += ws[(size_t)b * d_pad + j];
        sm[y][x] = s;
        __syncthreads();
        if (y == 0 && j < d_pad) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < REDUCE_SLICES; ++q) t += sm[q][x];
            out[j] = t;
        }
    }
    if (blockIdx.x == 0 && x == 0 && y == 0) {
        double f = 0.0;
        for (int b = 0; b < G; ++b) f = fmax_mode ? fmax(f, fws[b]) : f + fws[b];
        *fout = f;
    }
}
static inline void launch_reduce_ws(ciao_ctx *c, const double *ws, const double *fws, int G, double *out, double *fout, int fmax_mode,
                                    int with_vec) {
    reduce_ws_kernel<<<(int)((c->d_pad + 31) / 32), dim3(32, REDUCE_SLICES), 0, c->stream>>>(ws, fws, G, c->d_pad, out, fout, fmax_mode,
                                                                                          with_vec);
}

// ---------------------------------------------------------------------------
// The tail of a pass, ONE kernel: fixed-order sum of the CTA partials → (G ranks: one-shot exchange over peer memory, sum in
// rank order) → out = base + scale·(Σ/den).  Replaces reduce_ws_kernel → ncclAllReduce → finish_kernel (round 1: 0.265 ms
// per pass at 8 GPUs, ten times what a 32 KiB exchange should cost).
//   block = 32 columns × REDUCE_SLICES slices; column j < d_pad comes from ws[b][j], column j == d_pad is the scalar
//   (Σ_b fws[b], or max_b for the row-norm pass).  Exchange, per 32-column chunk and independent of all other chunks:
//   warp r < world stores the chunk into rank r's mail slot [seq & 1][my rank] (NVLink / local), one thread per peer
//   fences (system scope) and raises flag[my rank][chunk] = seq there, then waits for flag[r][chunk] ≥ seq in its own
//   arena (bounded wait: a missing peer surfaces as CIAO_ERR_COMM), and the chunk is summed over the slots 0 … world − 1 in
//   rank order — bit-identical on every rank, no floating-point atomics, no dependence on arrival order.
struct TailArgs {
    const double *ws, *fws;      // [G][d_pad], [G] (fws may be null: no scalar column)
    int G, len, fmax_mode, with_vec, chunk0;
    int64_t d_pad;
    int world, rank;
    uint32_t seq;
    uint64_t timeout_ns;
    double *arena[CIAO_MAX_PEERS];
    double *sum_out;             // [len]: the raw all-rank sum (vector, then the scalar)
    const double *base;          // FIN_PASS: out = base + scale·(Σ/den)   (out == nullptr: none)
    double scale, den;
    double *out;
    // closing update of a minibatch whose rows are sharded over the ranks (batch.cu run_batch_step_sharded):
    //   FIN_FINITO   av += Σ;  z = prox_g(av, γ̂)                                   Finito_basic.jl:115-118
    //   FIN_LFINITO  av += Σ + (Σ_i γ̂/γ_i)·(z − z_full);  z = prox_g(av, γ̂) unless last    Finito_LFinito.jl:92-98
    int fin_mode, last_batch;
    double *av, *z;
    const double *zf;
    double hat_gamma;
    RegParams reg;
    int *err;
};
enum { FIN_PASS = 0, FIN_FINITO = 1, FIN_LFINITO = 2 };

__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(32 * REDUCE_SLICES) pass_tail_kernel(const TailArgs a) {
    __shared__ double sm[REDUCE_SLICES][33];
    const int x = threadIdx.x, y = threadIdx.y;
    const int chunk = blockIdx.x + a.chunk0;
    const int64_t j = chunk * 32ll + x;
    const bool scalar_col = a.fws != nullptr && j == a.d_pad;
    const bool is_max = scalar_col && a.fmax_mode;
    double s = 0.0;
    if (j < a.d_pad) {
        if (a.with_vec) {   // four independent chains: the loads of a slice are in flight together (fixed order → deterministic)
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int b = y;
            for (; b + 3 * REDUCE_SLICES < a.G; b += 4 * REDUCE_SLICES) {
                s0 += a.ws[(size_t)b * a.d_pad + j];
                s1 += a.ws[(size_t)(b + REDUCE_SLICES) * a.d_pad + j];
                s2 += a.ws[(size_t)(b + 2 * REDUCE_SLICES) * a.d_pad + j];
                s3 += a.ws[(size_t)(b + 3 * REDUCE_SLICES) * a.d_pad + j];
            }
            for (; b < a.G; b += REDUCE_SLICES) s0 += a.ws[(size_t)b * a.d_pad + j];
            s = (s0 + s1) + (s2 + s3);
        }
    } else if (scalar_col) {
        for (int b = y; b < a.G; b += REDUCE_SLICES) s = a.fmax_mode ? fmax(s, a.fws[b]) : s + a.fws[b];
    }
    sm[y][x] = s;
    __syncthreads();
    double t = 0.0;
    if (y == 0) {
#pragma unroll
        for (int q = 0; q < REDUCE_SLICES; ++q) t = is_max ? fmax(t, sm[q][x]) : t + sm[q][x];
    }
    if (a.world > 1) {
        const int buf = (int)(a.seq & 1u);
        __syncthreads();
        if (y == 0) sm[0][x] = t;
        __syncthreads();
        if (y < a.world && j < a.len)    // warp y → rank y's arena, slot [buf][my rank], 256 contiguous bytes
            st_relaxed_sys_f64(a.arena[y] + ((size_t)buf * CIAO_MAX_PEERS + a.rank) * P2P_CAP + j, sm[0][x]);
        __syncwarp();
        if (x == 0 && y < a.world) {
            uint32_t *flags_there = reinterpret_cast<uint32_t *>(a.arena[y] + P2P_MAIL_DOUBLES);
            __threadfence_system();       // the warp's stores (ordered before this thread by the warp barrier) before the flag
            st_release_sys_u32(flags_there + a.rank * P2P_CHUNKS + chunk, a.seq);
            const uint32_t *flag_here = reinterpret_cast<const uint32_t *>(a.arena[a.rank] + P2P_MAIL_DOUBLES) + y * P2P_CHUNKS + chunk;
            const uint64_t t0 = global_timer_ns();
            while ((int32_t)(ld_acquire_sys_u32(flag_here) - a.seq) < 0) {
                if (global_timer_ns() - t0 > a.timeout_ns) {
                    atomicExch(a.err, 3);
                    break;
                }
                __nanosleep(100);
            }
        }
        __syncthreads();
        if (y == 0 && j < a.len) {
            const double *mine = a.arena[a.rank] + (size_t)buf * CIAO_MAX_PEERS * P2P_CAP + j;
            t = 0.0;
            for (int r = 0; r < a.world; ++r) {
                const double v = ld_relaxed_sys_f64(mine + (size_t)r * P2P_CAP);
                t = is_max ? fmax(t, v) : t + v;
            }
        }
    }
    if (a.fin_mode != FIN_PASS) {   // minibatch closing update on this chunk's columns
        // LFinito needs the batch's scalar Σ_i γ̂/γ_i in every chunk: each block sums the CTA partials itself (G ≤ a few hundred
        // L2-resident doubles) and, with several ranks, waits for the scalar chunk's flags and adds the ranks' slots in rank order
        __shared__ double fs_sh;
        if (a.fin_mode == FIN_LFINITO) {
            double f = 0.0;
            if (y == 0 && x == 0) {
                if (a.world > 1) {
                    const int sc = (int)(a.d_pad / 32);
                    const uint32_t *flags = reinterpret_cast<const uint32_t *>(a.arena[a.rank] + P2P_MAIL_DOUBLES);
                    const uint64_t t0 = global_timer_ns();
                    for (int r = 0; r < a.world; ++r)
                        while ((int32_t)(ld_acquire_sys_u32(flags + r * P2P_CHUNKS + sc) - a.seq) < 0) {
                            if (global_timer_ns() - t0 > a.timeout_ns) {
                                atomicExch(a.err, 3);
                                break;
                            }
                            __nanosleep(100);
                        }
                    const double *slot = a.arena[a.rank] + (size_t)(a.seq & 1u) * CIAO_MAX_PEERS * P2P_CAP + a.d_pad;
                    for (int r = 0; r < a.world; ++r) f += ld_relaxed_sys_f64(slot + (size_t)r * P2P_CAP);
                } else {
                    // same slice order as the scalar column above, so every block and the scalar chunk agree bit for bit
                    for (int q = 0; q < REDUCE_SLICES; ++q) {
                        double sq = 0.0;
                        for (int b = q; b < a.G; b += REDUCE_SLICES) sq += a.fws[b];
                        f += sq;
                    }
                }
                fs_sh = f;
            }
            __syncthreads();
        }
        if (y == 0 && j < a.d_pad) {
            double anew = __dadd_rn(a.av[j], t);
            if (a.fin_mode == FIN_LFINITO) anew = __dadd_rn(anew, __dmul_rn(fs_sh, __dsub_rn(a.z[j], a.zf[j])));
            a.av[j] = anew;
            if (a.fin_mode == FIN_FINITO || !a.last_batch) {
                const double lo = a.reg.lo_v ? a.reg.lo_v[j] : a.reg.lo_s, hi = a.reg.hi_v ? a.reg.hi_v[j] : a.reg.hi_s;
                a.z[j] = prox_rt(a.reg.kind, anew, a.hat_gamma * a.reg.lambda, lo, hi);
            }
        }
        return;
    }
    if (y == 0 && j < a.len) {
        a.sum_out[j] = t;
        if (a.out && j < a.d_pad) {
            double v = t;
            if (a.den != 1.0) v = __ddiv_rn(v, a.den);
            if (a.scale != 1.0) v = __dmul_rn(a.scale, v);
            a.out[j] = a.base ? __dadd_rn(a.base[j], v) : v;
        }
    }
}

// out = base + scale·(partial/den)   (the NCCL route and the non-row passes; the row passes fuse it into pass_tail_kernel)
__global__ void finish_kernel(const double *partial, const double *base, double scale, double den, int64_t d_pad, double *out) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= d_pad) return;
    double v = partial[j];
    if (den != 1.0) v = __ddiv_rn(v, den);
    if (scale != 1.0) v = __dmul_rn(scale, v);
    out[j] = base ? __dadd_rn(base[j], v) : v;
}
struct FinishSpec {
    const double *base;
    double scale, den;
    double *out;
};
static int run_finish(ciao_ctx *c, const double *base, double scale, double den, double *out) {
    finish_kernel<<<(int)((c->d_pad + 255) / 256), 256, 0, c->stream>>>(c->partial, base, scale, den, c->d_pad, out);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

static void fill_exchange(ciao_ctx *c, TailArgs &t, bool exchange) {
    t.world = exchange ? c->world : 1;
    t.rank = c->rank;
    t.seq = 0;
    t.timeout_ns = c->p2p_timeout_ns;
    for (int r = 0; r < CIAO_MAX_PEERS; ++r) t.arena[r] = c->p2p_peer[r];
    t.err = c->err_dev;
    if (exchange) t.seq = ++c->p2p_seq;
}

// in-place all-reduce of a plain device buffer through the peer exchange (collective; count ≤ P2P_CAP)
int run_p2p_allreduce(ciao_ctx *c, double *buf, int64_t count, int op_max) {
    if (count < 1 || count > P2P_CAP || (op_max && count != 1))
        CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "peer exchange: 1..%d doubles per collective (max: one scalar)", P2P_CAP);
    TailArgs t;
    memset(&t, 0, sizeof(t));
    if (op_max) {   // the scalar column carries the maximum
        t.ws = nullptr; t.fws = buf; t.d_pad = 0; t.with_vec = 0; t.fmax_mode = 1;
    } else {
        t.ws = buf; t.fws = nullptr; t.d_pad = count; t.with_vec = 1; t.fmax_mode = 0;
    }
    t.G = 1; t.len = (int)count; t.chunk0 = 0;
    t.sum_out = buf; t.out = nullptr;
    fill_exchange(c, t, true);
    pass_tail_kernel<<<(int)((count + 31) / 32), dim3(32, REDUCE_SLICES), 0, c->stream>>>(t);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
template <int CPT, int MODE, int LOSS>
static int launch_one(ciao_ctx *c, const PassArgs &a, int grid, int T, size_t smem) {
    auto kern = row_pass_kernel<CPT, MODE, LOSS>;
    static size_t configured[CIAO_MAX_DEVICES] = {};  // per device: the attribute is per device, and setting it costs several µs
    if (smem > configured[c->device % CIAO_MAX_DEVICES]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[c->device % CIAO_MAX_DEVICES] = smem;
    }
    kern<<<grid, T, smem, c->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return CIAO_OK;
}

template <int CPT>
static int launch_cpt(ciao_ctx *c, int mode, const PassArgs &a, int grid, int T, size_t smem) {
    const bool ls = c->loss_kind == CIAO_LOSS_LS;
    switch (mode) {
        case PASS_GRAD:
            return ls ? launch_one<CPT, PASS_GRAD, CIAO_LOSS_LS>(c, a, grid, T, smem)
                      : launch_one<CPT, PASS_GRAD, CIAO_LOSS_LOGISTIC>(c, a, grid, T, smem);
        case PASS_SAGA_INIT:
            return ls ? launch_one<CPT, PASS_SAGA_INIT, CIAO_LOSS_LS>(c, a, grid, T, smem)
                      : launch_one<CPT, PASS_SAGA_INIT, CIAO_LOSS_LOGISTIC>(c, a, grid, T, smem);
        case PASS_FINITO_INIT:
            return ls ? launch_one<CPT, PASS_FINITO_INIT, CIAO_LOSS_LS>(c, a, grid, T, smem)
                      : launch_one<CPT, PASS_FINITO_INIT, CIAO_LOSS_LOGISTIC>(c, a, grid, T, smem);
        default:
            return launch_one<CPT, PASS_NORMS, CIAO_LOSS_LS>(c, a, grid, T, smem);
    }
}

int ciao_comm_allreduce(ciao_ctx *c, double *buf, int64_t count, int op_max);  // comm.cu
int run_block_table_init(ciao_ctx *c, int mode, const double *x0_dev, int *grid_out);   // blockseq.cu
int ciao_comm_allgather_inplace(ciao_ctx *c, double *buf, int64_t count_per_rank);

// Runs one streaming pass.  Result: c->partial[0..d_pad) = Σ (unscaled, all ranks), c->partial[d_pad] = Σ f_i (or max).
// With `fin` the closing update out = base + scale·(Σ/den) rides in the same tail kernel.
int run_row_pass(ciao_ctx *c, int mode, const double *x_dev, bool cache_cz = false, const FinishSpec *fin = nullptr) {
    NvtxRange nvtx(mode == PASS_GRAD ? "ciao:pass:full_gradient" : mode == PASS_NORMS ? "ciao:pass:row_norms" : "ciao:pass:table_init");
    if (c->loss_kind != CIAO_LOSS_LS && c->loss_kind != CIAO_LOSS_LOGISTIC)
        CIAO_FAIL(CIAO_ERR_STATE, "row pass: no row problem set (ciao_set_rows / ciao_gen_synthetic first)");
    const int64_t d_pad = c->d_pad;
    const bool init_mode_blk = mode == PASS_SAGA_INIT || mode == PASS_FINITO_INIT;
    if (c->M > 1 && (c->world > 1 || c->win_n > 0)) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "block components are not sharded or windowed");
    if (init_mode_blk && c->M > 1) {   // per-component table rows: the general block kernel (blockseq.cu), closed by the same tail kernel
        int grid_b = 0;
        CIAO_TRY(run_block_table_init(c, mode, x_dev, &grid_b));
        TailArgs tb;
        memset(&tb, 0, sizeof(tb));
        tb.ws = c->ws; tb.fws = c->ws + (size_t)grid_b * d_pad; tb.G = grid_b; tb.d_pad = d_pad; tb.len = (int)d_pad + 1;
        tb.with_vec = 1; tb.sum_out = c->partial;
        if (fin) { tb.base = fin->base; tb.scale = fin->scale; tb.den = fin->den; tb.out = fin->out; }
        fill_exchange(c, tb, false);
        pass_tail_kernel<<<(int)((d_pad + 1 + 31) / 32), dim3(32, REDUCE_SLICES), 0, c->stream>>>(tb);
        CUDA_TRY(cudaGetLastError());
        c->timing.launches += 1;
        return CIAO_OK;
    }
    // Launch shape (scripts/k2_tune.py, gpurun_out/k2_tune.log: sweeps at d = 512 … 4096, 8.6 GB per pass).  A row should be one
    // group (16 columns per thread, one block reduction per 16 elements of work) and an SM should hold ≈ 512 threads in as many
    // small CTAs as that takes, each with a 2-stage ring: d = 1024 runs at 6.3 TB/s with 64 threads × 8 CTAs/SM against 3.1 TB/s
    // with 256 × 2; d = 4096 is 256 × 2 either way (7.0 vs 6.7 TB/s with 2 instead of 3 stages).  The table-init modes also
    // write N×d and prefer twice the threads per SM in CTAs of d/8 threads.
    const bool init_mode = mode == PASS_SAGA_INIT || mode == PASS_FINITO_INIT;
    int T_target = c->pass_threads > 0 ? c->pass_threads : (init_mode ? (int)std::min<int64_t>(256, std::max<int64_t>(64, d_pad / 8)) : 64);
    int cpt = 2;
    while (cpt < 16 && (d_pad + cpt - 1) / cpt > T_target) cpt *= 2;
    int64_t Tn = ((d_pad + cpt - 1) / cpt + 31) / 32 * 32;
    if (Tn > 512) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "row pass: d = %lld exceeds the engine limit 8192", (long long)c->d);
    const int T = (int)Tn;
    const int rpg = 16 / cpt;
    const size_t stage_bytes = (size_t)rpg * c->ld * sizeof(double);
    const int ctas_per_sm = c->pass_ctas > 0 ? c->pass_ctas : std::max(1, std::min(8, (init_mode ? 1024 : 512) / T));
    const size_t fixed = 2 * rpg * 32 * sizeof(double) + 16 * sizeof(uint64_t) + 256;
    const size_t budget = (size_t)(227 * 1024) / ctas_per_sm - (ctas_per_sm > 1 ? 1024 : 0);
    int S = c->pass_stages > 0 ? c->pass_stages : (ctas_per_sm > 1 ? 2 : 3);
    while (S > 1 && (size_t)S * stage_bytes + fixed > budget) --S;
    if ((size_t)S * stage_bytes + fixed > budget) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "row pass: stage does not fit shared memory");
    if (S > 16) S = 16;
    const size_t smem = (size_t)S * stage_bytes + fixed;
    const bool windowed = c->win_n > 0 && (mode == PASS_GRAD || mode == PASS_NORMS);
    const int64_t w0 = windowed ? c->win0 : 0, wn = windowed ? c->win_n : c->n_rows * c->M;   // rows streamed (M per component)
    const int64_t n_groups = (wn + rpg - 1) / rpg;
    int grid = (int)std::min<int64_t>(n_groups, (int64_t)c->num_sms * ctas_per_sm);
    if (grid < 1) grid = 1;

    const size_t need = ((size_t)grid * d_pad + grid + 16) * sizeof(double);
    if (need > c->ws_bytes) {
        if (c->ws) cudaFree(c->ws);
        c->ws = nullptr;
        CUDA_TRY(cudaMalloc(&c->ws, need));
        c->ws_bytes = need;
    }
    PassArgs a;
    a.rec = c->rec + w0 * c->ld; a.n_rows = wn; a.ld = c->ld;
    a.ss_out = nullptr;
    // the per-row step scalars for the inner kernels: all rows in one process, or — replicated rows, uniformly windowed
    // passes — each rank its window, all-gathered after the kernel
    const bool can_gather = c->world > 1 && c->nccl_comm != nullptr;   // the all-gather of the step scalars is NCCL's
    const bool gather_win = cache_cz && mode == PASS_GRAD && windowed && can_gather && c->win_uniform;
    // … or row-sharded data with the peers' shards attached (ciao_attach_peer_rows): every rank knows all shard bounds, so
    // all ranks take the same decision; the array is indexed by the global row
    bool gather_shard = cache_cz && mode == PASS_GRAD && !windowed && can_gather && c->peers.n == c->world &&
                        c->N_total % c->world == 0 && c->n_rows == c->N_total / c->world;
    for (int sidx = 0; gather_shard && sidx < c->peers.n; ++sidx)
        gather_shard = c->peers.start[sidx] == (int64_t)sidx * c->n_rows;
    gather_shard = gather_shard && c->row0 == (int64_t)c->rank * c->n_rows;
    const bool gather_ss = gather_win || gather_shard;
    // (a rank of a sharded problem without gathered scalars still keeps the scalars of ITS rows: the LFinito minibatch kernel
    // reads c_i(z_full) of the local rows from there instead of forming a second dot product)
    const bool local_only = !windowed && c->world > 1 && !gather_ss;
    if (cache_cz && c->M == 1 && mode == PASS_GRAD && ((!windowed && c->world == 1) || gather_ss || local_only)) {
        const int64_t ss_rows = gather_shard ? c->N_total : c->n_rows;
        if (c->ss && c->ss_cap < ss_rows) {
            cudaFree(c->ss);
            c->ss = nullptr;
        }
        if (!c->ss) {
            CUDA_TRY(cudaMalloc(&c->ss, (size_t)(ss_rows + 1) * 4 * sizeof(double)));
            c->ss_cap = ss_rows;
        }
        a.ss_out = c->ss + 4 * (gather_shard ? c->row0 : w0);
    }
    if (mode == PASS_GRAD && cache_cz) {
        c->cz_valid = a.ss_out != nullptr && !local_only;          // the sequential kernels need the scalars of ALL rows
        c->cz_local_valid = a.ss_out != nullptr && !windowed;
        c->ss_row0 = gather_shard ? c->row0 : 0;
    }
    a.d_pad = d_pad; a.x = x_dev;
    a.ws = c->ws; a.fws = c->ws + (size_t)grid * d_pad; a.table = c->table;
    a.Nd = (double)c->N_total; a.stages = S;
    if ((mode == PASS_SAGA_INIT || mode == PASS_FINITO_INIT) && !c->table)
        CIAO_FAIL(CIAO_ERR_STATE, "table init pass without a table");

    CUDA_TRY(cudaEventRecord(c->ev_pa, c->stream));
    int rc;
    switch (cpt) {
        case 2: rc = launch_cpt<2>(c, mode, a, grid, T, smem); break;
        case 4: rc = launch_cpt<4>(c, mode, a, grid, T, smem); break;
        case 8: rc = launch_cpt<8>(c, mode, a, grid, T, smem); break;
        default: rc = launch_cpt<16>(c, mode, a, grid, T, smem); break;
    }
    CIAO_TRY(rc);
    CUDA_TRY(cudaEventRecord(c->ev_pb, c->stream));
    // tail: CTA partials → (peer exchange) → closing update, one kernel; with NCCL only (no peer arenas attached) the exchange
    // and the closing update are separate launches
    const bool p2p = c->world > 1 && c->p2p_ready;
    const bool nccl_route = c->world > 1 && !p2p;
    TailArgs t;
    memset(&t, 0, sizeof(t));
    t.ws = a.ws; t.fws = a.fws; t.G = grid; t.d_pad = d_pad; t.len = (int)d_pad + 1;
    t.fmax_mode = mode == PASS_NORMS; t.with_vec = mode != PASS_NORMS;
    t.chunk0 = mode == PASS_NORMS ? (int)(d_pad / 32) : 0;
    t.sum_out = c->partial;
    if (fin && !nccl_route) {
        t.base = fin->base; t.scale = fin->scale; t.den = fin->den; t.out = fin->out;
    }
    fill_exchange(c, t, p2p);
    const int tail_grid = mode == PASS_NORMS ? 1 : (int)((d_pad + 1 + 31) / 32);
    pass_tail_kernel<<<tail_grid, dim3(32, REDUCE_SLICES), 0, c->stream>>>(t);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev_pc, c->stream));
    c->timing.launches += 2;
    c->pass_timed = true;
    c->tail_timed = true;
    c->timing.last_pass_bytes = (int64_t)wn * c->ld * 8 +
                                ((mode == PASS_SAGA_INIT || mode == PASS_FINITO_INIT) ? (int64_t)c->n_rows * d_pad * 8 : 0);
    if (nccl_route) {
        if (mode == PASS_NORMS) {
            CIAO_TRY(ciao_comm_allreduce(c, c->partial + d_pad, 1, 1));
        } else {
            CIAO_TRY(ciao_comm_allreduce(c, c->partial, d_pad + 1, 0));
        }
        if (fin) CIAO_TRY(run_finish(c, fin->base, fin->scale, fin->den, fin->out));
    }
    if (gather_ss) CIAO_TRY(ciao_comm_allgather_inplace(c, c->ss, 4 * wn));
    return CIAO_OK;
}

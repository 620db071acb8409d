// proshi.cu — K7: ProShI block updates for the sharing problem
//   (1/N) Σ f_i(x_i) + g(Σ x_i),   f_i = ½x'diag(q_i)x + c_i'x + (η/2)dist²(x,[lo,hi])   (test_sharing.jl:15-22)
//
//   proshi_init_kernel      s_i = x0 − (γ_i/N)∇f_i(x0);  Σ s_i           ProShI_basic.jl:76-83
//   proshi_steps_kernel     block steps + dual update                     ProShI_basic.jl:111-123
//   proshi_solution_kernel  s_i += γ_i z  IN PLACE                        ProShI_basic.jl:127-132
//
// Design.  With a diagonal Q_i the block update has no coupling
// between coordinates: every column j of (s, av, z) evolves independently given the
// index sequence.  proshi_steps_kernel (batch 1) keeps z_j and av_j of a thread's columns in
// registers for the whole call while producer lanes TMA-stage the slices of (q_i, c_i) and (γ_i, γ_i/N)
// and a table producer warp cp.async-stages the slice of s_i sixteen steps ahead — no barrier, no reduction, no kernel launch per
// step; proshi_batch_kernel (batches ≥ 64 blocks) works through the blocks of a batch in
// parallel, four lanes per block, and closes each batch with a fixed-order CTA reduction.
#include <algorithm>

#include "common.cuh"

struct ProshiArgs {
    const double *qd, *ql;  // [N][n_pad]
    double *table;          // [N][n_pad]
    const double *gam;      // [N]
    const double *gpair;    // [N][2]  (γ_i, γ_i/N): the pair a step needs, 16 bytes (γ_i/N precomputed, ProShI_basic.jl:116)
    const int64_t *idx;     // prepared
    int64_t K, N, n_pad;
    double *v_z, *v_av;
    double box_lo, box_hi, eta, Nd, hat_gamma;
    RegParams reg;
    const int *err;         // context error flag (out-of-range index seen by prep_indices_kernel): no block step runs
};

__device__ __forceinline__ double proshi_grad(double q, double c, double s, double lo, double hi, double eta) {
    // Sum(Quadratic, SqrDistL2): (q·s + c) + η·(s − Π_box s)
    const double g1 = __dadd_rn(__dmul_rn(q, s), c);
    const double pr = s < lo ? lo : (s > hi ? hi : s);
    const double g2 = __dmul_rn(eta, __dsub_rn(s, pr));
    return __dadd_rn(g1, g2);
}


// Batch-1 steps: a true dependency chain per column (z_j → s_ij → av_j → z_j), so the kernel is latency/issue bound and
// everything that is not on that chain has to stay off the compute warp.  A CTA owns 32·CPT columns: ONE compute warp keeps
// z_j, av_j of its columns in registers for the whole call; TWO producer lanes (warps 1 and 2, even and odd steps) stage the
// column slices of (q_i, c_i), the pair (γ_i, γ_i/N) and the index word by TMA, a table producer warp the slice of s_i by
// cp.async, all for the block needed PROSHI_D steps later, into a full/empty mbarrier ring.  No reduction, no CTA barrier, no kernel launch per step.
// History: register prefetch with ld.global ran at 0.74 µs/block (ptxas tracks all those loads with one scoreboard, so the
// first use of step k's registers also waited for the load just issued for step k+D: one DRAM latency per step); per-thread
// cp.async groups at 0.30 µs/block (≈ 190 instructions per step on the one warp that also walks the chain); with the staging
// moved to producer lanes the compute warp issues ≈ 70.
constexpr int PROSHI_D = 16;  // ring depth; a staged table slice is stale if the block recurs within D + 1 steps → HAZARD flag
static_assert((PROSHI_D & (PROSHI_D - 1)) == 0 && PROSHI_D + 2 <= CIAO_HAZARD_WINDOW,
              "prep_indices_kernel must flag repeats within the prefetch window");

#ifndef PROSHI_NP
#define PROSHI_NP 2
#endif
constexpr int PROSHI_PRODUCERS = PROSHI_NP;  // producer lanes (one warp each), step st is staged by producer st mod NP

template <int CPT, int REG>
__global__ void __launch_bounds__(32 * (2 + PROSHI_PRODUCERS)) proshi_steps_kernel(const ProshiArgs p) {
    if (*reinterpret_cast<const volatile int *>(p.err) != 0) return;   // ProShI_basic.jl:113 would throw BoundsError first
    constexpr int D = PROSHI_D;
    constexpr int COLS = 32 * CPT;                 // columns per CTA
    constexpr int SLOT = 3 * COLS + 4;             // doubles: q | c | s | γ, γ/N | index word | pad
    __shared__ __align__(128) double ring[D * SLOT];
    __shared__ uint64_t full_bar[D], empty_bar[D];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t col0 = (int64_t)blockIdx.x * COLS;
    const int64_t ncol = min((int64_t)COLS, p.n_pad - col0);   // columns of this CTA that exist (multiple of 4)
    const int64_t K = p.K;
    for (int i = tid; i < D * SLOT; i += blockDim.x) ring[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < D; ++s) {
            mbar_init(&full_bar[s], 33);   // lane 0's arrive.expect_tx + one cp.async arrive per lane of the producer warp
            mbar_init(&empty_bar[s], 1);
        }
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();

    if (warp >= 1 && warp <= PROSHI_PRODUCERS) {
        // ===================== producer lanes: warp 1 stages the even steps, warp 2 the odd ones =====================
        // bulk copies (TMA) of the read-only slices q_i, c_i and of the pair (γ_i, γ_i/N), and the index word
        if (lane == 0) {
            const int Ki = (int)K;
            const uint32_t ring_s = smem_u32(ring), full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar);
            const uint32_t bytes = (uint32_t)ncol * 8;
            const uint32_t tx = 2 * bytes + 16;
            for (int st = warp - 1; st < Ki; st += PROSHI_PRODUCERS) {
                const uint32_t slot = (uint32_t)st & (D - 1);
                if (st >= D) mbar_wait_s(empty_s + slot * 8, (((uint32_t)st / D) - 1u) & 1u);  // step st − D has left the slot
                const int64_t pidx = __ldg(p.idx + st);
                const int64_t i = pidx & CIAO_IDX_MASK;
                const int64_t off = i * p.n_pad + col0;
                const uint32_t dst = ring_s + slot * (SLOT * 8), bar = full_s + slot * 8;
                sts_b64(dst + (3 * COLS + 2) * 8, pidx);  // released by the arrive below
                mbar_arrive_expect_tx_s(bar, tx);
                tma_load_1d_s(dst, p.qd + off, bytes, bar);
                tma_load_1d_s(dst + COLS * 8, p.ql + off, bytes, bar);
                tma_load_1d_s(dst + 3 * COLS * 8, p.gpair + 2 * i, 16, bar);
            }
        }
        return;
    }
    if (warp == PROSHI_PRODUCERS + 1) {
        // ===================== table producer warp =====================
        // The TABLE slice is written inside this kernel by the compute warp, so it is staged through the same (generic) proxy:
        // 16-byte cp.async copies by the lanes of this warp (see seq_impl.cuh for why not TMA).  The compute warp's st.global of
        // step k is ordered before them by its warp barrier + release arrive on empty_bar in take() and by the acquire wait of
        // every lane here.  Every repeat at distance ≥ PROSHI_D + 1 is ordered that way, closer ones carry the HAZARD flag.
        const int Ki = (int)K;
        const uint32_t ring_s = smem_u32(ring), full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar);
        const uint32_t n_chunks = (uint32_t)ncol / 2;   // 16-byte chunks of the table slice (≤ 32)
        for (int st = 0; st < Ki; ++st) {
            const uint32_t slot = (uint32_t)st & (D - 1);
            if (st >= D) mbar_wait_s(empty_s + slot * 8, (((uint32_t)st / D) - 1u) & 1u);
            const int64_t i = __ldg(p.idx + st) & CIAO_IDX_MASK;
            if ((uint32_t)lane < n_chunks)
                cp_async_16(ring_s + slot * (SLOT * 8) + 2 * COLS * 8 + lane * 16, p.table + i * p.n_pad + col0 + 2 * lane);
            cp_async_arrive_noinc(full_s + slot * 8);
        }
        return;
    }
    // ===================== compute warp =====================
    const int lc = CPT * lane;                      // first local column of this lane
    const bool active = lc < ncol;
    const int64_t col = col0 + (active ? lc : 0);
    double z[CPT], av[CPT], lo[CPT], hi[CPT];
#pragma unroll
    for (int e = 0; e < CPT; ++e) {
        z[e] = active ? p.v_z[col + e] : 0.0;
        av[e] = active ? p.v_av[col + e] : 0.0;
        lo[e] = (active && p.reg.lo_v) ? p.reg.lo_v[col + e] : p.reg.lo_s;
        hi[e] = (active && p.reg.hi_v) ? p.reg.hi_v[col + e] : p.reg.hi_s;
    }
    const double gl = p.hat_gamma * p.reg.lambda;
    const double rhat = __ddiv_rn(1.0, p.hat_gamma);
    struct Blk {
        double q[CPT], c[CPT], s[CPT];
        double gi, gn;
        int64_t ik;
    };
    // pulls the staged block of `step` into registers and hands the slot back to the producers
    auto take = [&](int64_t step, Blk &b) {
        const int slot = (int)(step & (D - 1));
        const double *sp = ring + slot * SLOT;
        if (CPT == 2) {
            const double2 vq = *reinterpret_cast<const double2 *>(sp + lc), vc = *reinterpret_cast<const double2 *>(sp + COLS + lc),
                          vs = *reinterpret_cast<const double2 *>(sp + 2 * COLS + lc);
            b.q[0] = vq.x; b.q[CPT - 1] = vq.y; b.c[0] = vc.x; b.c[CPT - 1] = vc.y; b.s[0] = vs.x; b.s[CPT - 1] = vs.y;
        } else {
            b.q[0] = sp[lc]; b.c[0] = sp[COLS + lc]; b.s[0] = sp[2 * COLS + lc];
        }
        const double2 g2 = *reinterpret_cast<const double2 *>(sp + 3 * COLS);
        b.gi = g2.x; b.gn = g2.y;
        b.ik = *reinterpret_cast<const int64_t *>(sp + 3 * COLS + 2);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
    };
    Blk cur;
    if (K > 0) {
        mbar_wait(&full_bar[0], 0);
        take(0, cur);
    }
    for (int64_t k = 0; k < K; ++k) {
        // is the next block staged?  asked now, looked at after this step's arithmetic (the test takes ≈ 150 cycles to answer)
        const int64_t k1 = k + 1;
        uint64_t *nbar = &full_bar[k1 & (D - 1)];
        const uint32_t npar = (uint32_t)((k1 / D) & 1);
        const uint32_t ready = (k1 < K) ? mbar_test(nbar, npar) : 1u;
        const int64_t ik = cur.ik;
        double *srow = p.table + (ik & CIAO_IDX_MASK) * p.n_pad + col;
        double s[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e) s[e] = cur.s[e];
        if (ik & CIAO_FLAG_HAZARD) {  // rewritten after its copy was staged: re-read behind our own store
#pragma unroll
            for (int e = 0; e < CPT; ++e) s[e] = __ldcg(srow + e);
        }
        const double gi = cur.gi, cneg = -cur.gn;
        double t[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e) {  // ProShI_basic.jl:113-119
            av[e] = __dsub_rn(av[e], s[e]);
            const double x = __dadd_rn(s[e], __dmul_rn(gi, z[e]));
            t[e] = __dadd_rn(__dmul_rn(proshi_grad(cur.q[e], cur.c[e], x, p.box_lo, p.box_hi, p.eta), cneg), x);
            av[e] = __dadd_rn(av[e], t[e]);
        }
        if (active) {
            if (CPT == 2) __stcg(reinterpret_cast<double2 *>(srow), make_double2(t[0], t[CPT - 1]));
            else __stcg(srow, t[0]);
        }
        if (ik & CIAO_FLAG_PROX) {  // :121-123
#pragma unroll
            for (int e = 0; e < CPT; ++e)
                z[e] = div_by(__dsub_rn(prox_elem<REG>(av[e], gl, lo[e], hi[e]), av[e]), p.hat_gamma, rhat);
        }
        if (k1 < K) {
            if (!ready) mbar_wait(nbar, npar);
            take(k1, cur);
        }
    }
    if (active) {
#pragma unroll
        for (int e = 0; e < CPT; ++e) {
            p.v_z[col + e] = z[e];
            p.v_av[col + e] = av[e];
        }
    }
}

// ---------------------------------------------------------------------------
// Minibatch path (batch ≥ PROSHI_BATCH_MIN blocks): inside a batch every block uses the same z
// (ProShI_basic.jl:111-120), so the blocks are independent and av only needs Σ_i (t_i − s_i).  A CTA owns
// 8 columns (one 64-byte chunk of every row) for the WHOLE call; FOUR LANES share a block — each loads one 16-byte piece
// of the chunk from each of the three arrays, so a warp-wide load instruction touches 8 blocks × 2 full sectors (the first
// version gave every lane its own block: 32 half-used sectors per instruction, 2.3 TB/s).  A fixed-order CTA
// reduction closes the batch and the dual update of the CTA's own columns follows — columns never interact, so there is
// no grid-wide synchronisation and the kernel streams the three block arrays.  Bitwise reproducible (fixed thread ↔ block
// mapping, fixed reduction order).
constexpr int PROSHI_BATCH_MIN = 64;
// PROSHI_BT threads per CTA = BT/4 block slots × 4 lanes, PROSHI_BU blocks per thread in flight.  Measured at N = 2^18, n = 1024
// (scripts/proshi_batch_probe.py): batch 4096 → 3.2 TB/s with (256, 4), 4.2 TB/s with (512, 4), 2.7 with (256, 8), 4.0 with
// (1024, 2); batch 256 → 2.3 TB/s with (256, 4), 2.1 with (512, 4).
template <int PROSHI_BT, int PROSHI_BU>
__global__ void __launch_bounds__(PROSHI_BT) proshi_batch_kernel(const ProshiArgs p, const int64_t *ptr, int64_t n_batches) {
    __shared__ double red[2][PROSHI_BT / 32][4][2];
    if (*reinterpret_cast<const volatile int *>(p.err) != 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = tid & 3, rs = tid >> 2;            // piece of the 64-byte chunk, block slot
    const int64_t col = 8 * (int64_t)blockIdx.x + 2 * h;
    const bool cv = col < p.n_pad;
    const int64_t colc = cv ? col : 0;
    double z0 = cv ? p.v_z[colc] : 0.0, z1 = cv ? p.v_z[colc + 1] : 0.0;
    double av0 = cv ? p.v_av[colc] : 0.0, av1 = cv ? p.v_av[colc + 1] : 0.0;
    const double lo0 = (cv && p.reg.lo_v) ? p.reg.lo_v[colc] : p.reg.lo_s, lo1 = (cv && p.reg.lo_v) ? p.reg.lo_v[colc + 1] : p.reg.lo_s;
    const double hi0 = (cv && p.reg.hi_v) ? p.reg.hi_v[colc] : p.reg.hi_s, hi1 = (cv && p.reg.hi_v) ? p.reg.hi_v[colc + 1] : p.reg.hi_s;
    const double gl = p.hat_gamma * p.reg.lambda;
    const double rhat = __ddiv_rn(1.0, p.hat_gamma);
    constexpr int SLOTS = PROSHI_BT / 4;

    for (int64_t b = 0; b < n_batches; ++b) {
        const int64_t lo_t = ptr[b], hi_t = ptr[b + 1];
        double ds0 = 0.0, ds1 = 0.0;
        for (int64_t t = lo_t + rs; t < hi_t; t += (int64_t)PROSHI_BU * SLOTS) {
            bool act[PROSHI_BU];
            int64_t off[PROSHI_BU];
            double gi[PROSHI_BU], cneg[PROSHI_BU];
            double2 q2[PROSHI_BU], c2[PROSHI_BU], s2[PROSHI_BU];
#pragma unroll
            for (int u = 0; u < PROSHI_BU; ++u) {     // all loads of the four blocks are in flight before the first use
                act[u] = cv && (t + (int64_t)u * SLOTS < hi_t);
                const int64_t i = act[u] ? (__ldg(p.idx + t + (int64_t)u * SLOTS) & CIAO_IDX_MASK) : 0;
                off[u] = i * p.n_pad + colc;
                const double2 g2 = __ldg(reinterpret_cast<const double2 *>(p.gpair) + i);
                gi[u] = g2.x;
                cneg[u] = -g2.y;
                q2[u] = __ldcs(reinterpret_cast<const double2 *>(p.qd + off[u]));
                c2[u] = __ldcs(reinterpret_cast<const double2 *>(p.ql + off[u]));
                s2[u] = __ldcg(reinterpret_cast<const double2 *>(p.table + off[u]));
            }
#pragma unroll
            for (int u = 0; u < PROSHI_BU; ++u) {
                if (!act[u]) continue;
                // ProShI_basic.jl:114-119 for two columns
                const double x0 = __dadd_rn(s2[u].x, __dmul_rn(gi[u], z0));
                const double x1 = __dadd_rn(s2[u].y, __dmul_rn(gi[u], z1));
                const double t0 = __dadd_rn(__dmul_rn(proshi_grad(q2[u].x, c2[u].x, x0, p.box_lo, p.box_hi, p.eta), cneg[u]), x0);
                const double t1 = __dadd_rn(__dmul_rn(proshi_grad(q2[u].y, c2[u].y, x1, p.box_lo, p.box_hi, p.eta), cneg[u]), x1);
                ds0 += __dsub_rn(t0, s2[u].x);
                ds1 += __dsub_rn(t1, s2[u].y);
                __stcg(reinterpret_cast<double2 *>(p.table + off[u]), make_double2(t0, t1));
            }
        }
        // close the batch: fixed-order reduction of Σ(t − s) over the block slots (lanes with the same piece h, then warps),
        // then av and the dual variable z (:121-123)
        const int par = (int)(b & 1);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            ds0 += __shfl_xor_sync(0xffffffffu, ds0, o);
            ds1 += __shfl_xor_sync(0xffffffffu, ds1, o);
        }
        if (lane < 4) {
            red[par][warp][lane][0] = ds0;
            red[par][warp][lane][1] = ds1;
        }
        __syncthreads();  // also orders this batch's table writes before the next batch's reads (same CTA, same columns)
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int w = 0; w < PROSHI_BT / 32; ++w) {
            s0 += red[par][w][h][0];
            s1 += red[par][w][h][1];
        }
        av0 = __dadd_rn(av0, s0);
        av1 = __dadd_rn(av1, s1);
        z0 = div_by(__dsub_rn(prox_rt(p.reg.kind, av0, gl, lo0, hi0), av0), p.hat_gamma, rhat);
        z1 = div_by(__dsub_rn(prox_rt(p.reg.kind, av1, gl, lo1, hi1), av1), p.hat_gamma, rhat);
    }
    if (tid < 4 && cv) {
        p.v_z[col] = z0; p.v_z[col + 1] = z1;
        p.v_av[col] = av0; p.v_av[col + 1] = av1;
    }
}

// grid (row groups, column chunks of 512); ws[blockIdx.x][n_pad] = partial Σ s_i
__global__ void __launch_bounds__(256) proshi_init_kernel(const double *qd, const double *ql, const double *gam,
                                                          double *gpair, const double *x0, double *table, double *ws, int64_t N,
                                                          int64_t n_pad, double box_lo, double box_hi, double eta,
                                                          double Nd) {
    const int64_t col = 2 * (blockIdx.y * (int64_t)blockDim.x + threadIdx.x);
    if (col >= n_pad) return;
    const double xa = x0[col], xb = x0[col + 1];
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = blockIdx.x; i < N; i += gridDim.x) {
        const double2 q = __ldcs(reinterpret_cast<const double2 *>(qd + i * n_pad + col));
        const double2 c = __ldcs(reinterpret_cast<const double2 *>(ql + i * n_pad + col));
        const double cg = __ddiv_rn(__ldg(gam + i), Nd);
        if (col == 0) reinterpret_cast<double2 *>(gpair)[i] = make_double2(__ldg(gam + i), cg);
        const double s0 = __dsub_rn(xa, __dmul_rn(cg, proshi_grad(q.x, c.x, xa, box_lo, box_hi, eta)));
        const double s1 = __dsub_rn(xb, __dmul_rn(cg, proshi_grad(q.y, c.y, xb, box_lo, box_hi, eta)));
        __stcs(reinterpret_cast<double2 *>(table + i * n_pad + col), make_double2(s0, s1));
        a0 += s0;
        a1 += s1;
    }
    ws[(size_t)blockIdx.x * n_pad + col] = a0;
    ws[(size_t)blockIdx.x * n_pad + col + 1] = a1;
}

// ws[blockIdx.x][n_pad] = partial Σ_i table_i   (sum(x_proshi), test_sharing.jl:42)
__global__ void __launch_bounds__(256) table_colsum_kernel(const double *table, double *ws, int64_t N, int64_t n_pad) {
    const int64_t col = 2 * (blockIdx.y * (int64_t)blockDim.x + threadIdx.x);
    if (col >= n_pad) return;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = blockIdx.x; i < N; i += gridDim.x) {
        const double2 s = __ldcs(reinterpret_cast<const double2 *>(table + i * n_pad + col));
        a0 += s.x;
        a1 += s.y;
    }
    ws[(size_t)blockIdx.x * n_pad + col] = a0;
    ws[(size_t)blockIdx.x * n_pad + col + 1] = a1;
}

// s_i += γ_i z for every block, in place (ProShI_basic.jl:127-132): a pure streaming read-modify-write of the table.  Four rows
// per thread are in flight (independent 128-bit loads issued before the first use; round 1's one-load-at-a-time loop ran at
// 2.9 TB/s of 16·N·n bytes), streaming cache hints on both sides.
__global__ void __launch_bounds__(256) proshi_solution_kernel(double *table, const double *gam, const double *z, int64_t N,
                                                              int64_t n_pad) {
    constexpr int U = 4;
    const int64_t col = 2 * (blockIdx.y * (int64_t)blockDim.x + threadIdx.x);
    if (col >= n_pad) return;
    const double z0 = z[col], z1 = z[col + 1];
    const int64_t stride = gridDim.x;
    int64_t i = blockIdx.x;
    for (; i + (U - 1) * stride < N; i += U * stride) {
        double2 s[U];
        double gi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s[u] = __ldcs(reinterpret_cast<const double2 *>(table + (i + u * stride) * n_pad + col));
            gi[u] = __ldg(gam + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s[u].x = __dadd_rn(s[u].x, __dmul_rn(gi[u], z0));
            s[u].y = __dadd_rn(s[u].y, __dmul_rn(gi[u], z1));
            __stcs(reinterpret_cast<double2 *>(table + (i + u * stride) * n_pad + col), s[u]);
        }
    }
    for (; i < N; i += stride) {
        double2 *sp = reinterpret_cast<double2 *>(table + i * n_pad + col);
        double2 s = __ldcs(sp);
        const double gi = __ldg(gam + i);
        s.x = __dadd_rn(s.x, __dmul_rn(gi, z0));
        s.y = __dadd_rn(s.y, __dmul_rn(gi, z1));
        __stcs(sp, s);
    }
}

// z = (prox_g(av, γ̂) − av)/γ̂     ProShI_basic.jl:84-86
__global__ void proshi_dual_kernel(const double *av, double *z, int64_t n_pad, double hat_gamma, RegParams reg) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    const double lo = reg.lo_v ? reg.lo_v[j] : reg.lo_s, hi = reg.hi_v ? reg.hi_v[j] : reg.hi_s;
    const double a = av[j];
    z[j] = __ddiv_rn(__dsub_rn(prox_rt(reg.kind, a, hat_gamma * reg.lambda, lo, hi), a), hat_gamma);
}

static int ws_reserve(ciao_ctx *c, size_t need) {
    if (need > c->ws_bytes) {
        if (c->ws) cudaFree(c->ws);
        c->ws = nullptr;
        c->ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->ws, need));
        c->ws_bytes = need;
    }
    return CIAO_OK;
}

static void grid2d(ciao_ctx *c, int64_t N, int64_t n_pad, dim3 *grid, int *G) {
    const int chunks = (int)((n_pad / 2 + 255) / 256);
    int rows = std::max(1, c->num_sms * 8 / chunks);
    if (rows > N) rows = (int)N;
    *grid = dim3(rows, chunks);
    *G = rows;
}

// leaves Σ (column sums) in c->partial[0..n_pad)
static int reduce_partials(ciao_ctx *c, int G) {
    launch_reduce_ws(c, c->ws, c->ws, G, c->partial, c->partial + c->d_pad, 0, 1);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

int run_proshi_init(ciao_ctx *c, const double *x0_dev) {
    dim3 grid; int G;
    grid2d(c, c->N_total, c->d_pad, &grid, &G);
    CIAO_TRY(ws_reserve(c, ((size_t)G * c->d_pad + 16) * sizeof(double)));
    if (!c->gpair) CUDA_TRY(cudaMalloc(&c->gpair, (size_t)c->N_total * 2 * sizeof(double)));
    CUDA_TRY(cudaEventRecord(c->ev_pa, c->stream));
    proshi_init_kernel<<<grid, 256, 0, c->stream>>>(c->qd, c->ql, c->gamma_dev, c->gpair, x0_dev, c->table, c->ws, c->N_total,
                                                    c->d_pad, c->box_lo, c->box_hi, c->eta, (double)c->N_total);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev_pb, c->stream));
    c->timing.launches += 1;
    c->timing.last_pass_bytes = c->N_total * c->d_pad * 24;
    c->pass_timed = true;
    return reduce_partials(c, G);
}

// gpair[i] = (γ_i, γ_i/N) without touching the table (ciao_solver_restore)
__global__ void proshi_gpair_kernel(const double *gam, double *gpair, int64_t N, double Nd) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) reinterpret_cast<double2 *>(gpair)[i] = make_double2(gam[i], __ddiv_rn(gam[i], Nd));
}
int run_proshi_gpair(ciao_ctx *c) {
    if (!c->gpair) CUDA_TRY(cudaMalloc(&c->gpair, (size_t)c->N_total * 2 * sizeof(double)));
    proshi_gpair_kernel<<<(int)((c->N_total + 255) / 256), 256, 0, c->stream>>>(c->gamma_dev, c->gpair, c->N_total, (double)c->N_total);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

int run_proshi_dual(ciao_ctx *c) {
    const int nb = (int)((c->d_pad + 255) / 256);
    proshi_dual_kernel<<<nb, 256, 0, c->stream>>>(ctx_vec(c, CIAO_VEC_AV), ctx_vec(c, CIAO_VEC_Z), c->d_pad, c->hat_gamma, c->reg);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

int run_proshi_steps(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, const int64_t *ptr_dev, int64_t n_batches) {
    NvtxRange nvtx("ciao:proshi:steps");
    if (K <= 0) return CIAO_OK;
    if (K >= (int64_t)1 << 31) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "ProShI: more than 2^31 - 1 block steps in one call");
    ProshiArgs a;
    a.qd = c->qd; a.ql = c->ql; a.table = c->table; a.gam = c->gamma_dev; a.gpair = c->gpair; a.idx = idx_prepared;
    a.K = K; a.N = c->N_total; a.n_pad = c->d_pad;
    a.v_z = ctx_vec(c, CIAO_VEC_Z); a.v_av = ctx_vec(c, CIAO_VEC_AV);
    a.box_lo = c->box_lo; a.box_hi = c->box_hi; a.eta = c->eta; a.Nd = (double)c->N_total; a.hat_gamma = c->hat_gamma;
    a.reg = c->reg;
    a.err = c->err_dev;
    CUDA_TRY(cudaEventRecord(c->ev_sa, c->stream));
    if (n_batches > 0 && K / n_batches >= PROSHI_BATCH_MIN) {
        // minibatches: blocks of a batch in parallel, one CTA per 8 columns
        if (K / n_batches >= 1024) proshi_batch_kernel<512, 4><<<(int)((c->d_pad + 7) / 8), 512, 0, c->stream>>>(a, ptr_dev, n_batches);
        else proshi_batch_kernel<256, 4><<<(int)((c->d_pad + 7) / 8), 256, 0, c->stream>>>(a, ptr_dev, n_batches);
    } else {
        // one compute warp + two producer lanes per CTA; one column per thread while that still gives ≤ 2 CTAs per SM
        const int64_t nc1 = (c->d_pad + 31) / 32;
        const bool one = nc1 <= 2 * (int64_t)c->num_sms;
        const int grid = (int)(one ? nc1 : (c->d_pad + 63) / 64);
        switch (c->reg.kind) {
            case CIAO_REG_NORML1:
                if (one) proshi_steps_kernel<1, CIAO_REG_NORML1><<<grid, 32 * (2 + PROSHI_PRODUCERS), 0, c->stream>>>(a);
                else proshi_steps_kernel<2, CIAO_REG_NORML1><<<grid, 32 * (2 + PROSHI_PRODUCERS), 0, c->stream>>>(a);
                break;
            case CIAO_REG_INDBOX:
                if (one) proshi_steps_kernel<1, CIAO_REG_INDBOX><<<grid, 32 * (2 + PROSHI_PRODUCERS), 0, c->stream>>>(a);
                else proshi_steps_kernel<2, CIAO_REG_INDBOX><<<grid, 32 * (2 + PROSHI_PRODUCERS), 0, c->stream>>>(a);
                break;
            default:
                if (one) proshi_steps_kernel<1, CIAO_REG_ZERO><<<grid, 32 * (2 + PROSHI_PRODUCERS), 0, c->stream>>>(a);
                else proshi_steps_kernel<2, CIAO_REG_ZERO><<<grid, 32 * (2 + PROSHI_PRODUCERS), 0, c->stream>>>(a);
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev_sb, c->stream));
    c->timing.launches += 1;
    c->timing.last_seq_steps = K;
    c->seq_timed = true;
    return CIAO_OK;
}

int run_proshi_solution(ciao_ctx *c) {
    dim3 grid; int G;
    grid2d(c, c->N_total, c->d_pad, &grid, &G);
    CUDA_TRY(cudaEventRecord(c->ev_pa, c->stream));
    proshi_solution_kernel<<<grid, 256, 0, c->stream>>>(c->table, c->gamma_dev, ctx_vec(c, CIAO_VEC_Z), c->N_total, c->d_pad);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev_pb, c->stream));
    c->timing.launches += 1;
    c->timing.last_pass_bytes = c->N_total * c->d_pad * 16;
    c->pass_timed = true;
    return CIAO_OK;
}

int run_table_colsum(ciao_ctx *c, int64_t n_rows) {
    dim3 grid; int G;
    grid2d(c, n_rows, c->d_pad, &grid, &G);
    CIAO_TRY(ws_reserve(c, ((size_t)G * c->d_pad + 16) * sizeof(double)));
    table_colsum_kernel<<<grid, 256, 0, c->stream>>>(c->table, c->ws, n_rows, c->d_pad);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return reduce_partials(c, G);
}

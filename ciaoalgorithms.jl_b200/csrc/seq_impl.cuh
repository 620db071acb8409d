// seq_impl.cuh — K3..K6: the sequential inner loops as ONE persistent thread-block-cluster
// kernel per call (no kernel launch per sample).
//
//   ALG_SVRG    SVRG_basic.jl:73-87      m inner steps + snapshot average
//   ALG_SAGA    SAGA_basic.jl:55-65      K single-sample steps with the N×d gradient table (SAGA and SAG)
//   ALG_FINITO  Finito_basic.jl:110-118  steps/batches with the N×d table s_i and per-component γ_i
//   ALG_LFINITO Finito_LFinito.jl:91-100 one sweep of corrections over all batches (no table)
//
// Design.  A cluster of C CTAs splits the d columns; a compute thread
// owns CPT fixed columns of every state vector (w/z, av, z_full, Σw) in REGISTERS for
// the whole call.  Each CTA has W compute warps, one PRODUCER warp and, for the table algorithms, one TABLE PRODUCER
// warp.  Per step:
//   1. producer lane: the sampled row slice a_i[cols of this CTA] + its record tail (b_i, λ_i,
//      γ_i/N, γ̂/γ_i, cached c_i(z_full)) + the prepared index word are staged D steps ahead into an
//      mbarrier ring by TMA (cp.async.bulk), driven by the host-generated index sequence.  A slot is free
//      once the exchange phase of the step that used it is over.  (SAGA/Finito) the same slice of the
//      table row s_i joins it in the slot — copied by the table producer warp with cp.async, the generic
//      proxy the rows are also written through (see the comment above CIAO table ordering below).
//   2. compute: partial dots → warp shuffles → each warp pushes its partial into EVERY CTA
//      of the cluster through DSMEM with st.async (a remote store that completes on the
//      destination CTA's mbarrier by tx-count, so neither side needs a cluster-scope
//      fence); every warp then reduces the C·W partials with the same tensor-core sum (two fp64
//      DMMAs + one add, no shuffle: common.cuh warp_sum_mma — a quarter of the latency of five shuffle rounds), so
//      all threads of all CTAs hold bit-identical scalars — no CTA or cluster barrier
//      instruction on the step's critical path.  While the exchange is in flight the row
//      and the table row of step k+1 are pulled into the other (ping-pong) register set.
//   3. the variance-reduced / aggregated update, the table row write and prox_g are
//      fused, element by element, in the reference's rounding order.
// Table rows are written by their owner threads only (st.global.cg → L2).  A staged copy is
// stale when the same row index occurs again within the D-step prefetch window; such steps are
// flagged by prep_indices_kernel (HAZARD) and re-read the row with ld.global.cg after the previous
// write (same thread wrote those addresses → coherent); every other repeat is ordered behind the write
// by a release/acquire chain through the "written" mbarrier.
// The step is latency/issue bound (one warp per SM sub-partition), so the loop body is
// kept free of predicates and global index loads: ring slots are zero padded to the thread
// grid and carry everything a step needs.
#pragma once
#include <algorithm>

#include "common.cuh"

struct SeqArgs {
    PeerTable rows;         // where row i lives (local HBM, or a peer's over NVLink)
    int64_t ld, d_pad, dc;  // dc = columns per CTA
    const int64_t *idx;     // prepared: 0-based row | flags
    int64_t K;
    double *table;
    const double *ss;       // [n_rows][4]: {b_i, λ_i, 0, c_i(z_full)} (CZ kernels), written by the last full-gradient pass
    double *v_z, *v_zfull, *v_w, *v_av, *v_zsum;
    double gamma, hat_gamma, Nd, m_d;
    int plus, sag;
    int npart_pad;          // C·W rounded up to a multiple of 32
    int table_tma;          // table rows are staged in the ring (cp.async by the producer warp; else: register prefetch with ld.global)
    int zero;               // always 0, opaque to the compiler (pins work in front of the exchange wait, see below)
    int active_cluster;     // the cluster of the grid that does the work (the others exit at once): cluster position = SM set
    long long *clk_out;     // [2]: SM cycles and nanoseconds (globaltimer) CTA 0 spent in the kernel → the SM clock it actually ran at
    int *smid_out;          // [C]: the SM each CTA of the cluster runs on (cross-SM DSMEM latency is a per-pair constant)
    const int *err;         // error flag of the context: set by prep_indices_kernel when an index is out of range → no step runs
    RegParams reg;
};

#ifdef CIAO_SEQ_PROFILE
static __device__ long long g_seq_prof[16 * 8];  // [CTA rank][phase]: accumulated cycles of thread 0 of every CTA (debug builds only)
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(i, a, b) prof_acc[i] += (b) - (a)
#else
#define PROF_T(var)
#define PROF_ADD(i, a, b)
#endif

constexpr int SEQ_D = 8;           // row ring depth (steps of prefetch), power of two
static_assert((SEQ_D & (SEQ_D - 1)) == 0, "ring slots and mbarrier parities are derived from step & (D - 1), step / D");
static_assert(SEQ_D + 1 <= CIAO_HAZARD_WINDOW - 1, "prep_indices_kernel must flag repeats within the prefetch window");
// Table rows (SAGA/Finito) are written by the compute threads with st.global.cg and re-read, D steps ahead of their use, into
// the ring.  Both sides stay in the GENERIC proxy: the staging copies are cp.async (LDGSTS, 16 bytes per lane of the producer
// warp, completion on the slot's mbarrier with cp.async.mbarrier.arrive), not TMA bulk copies — a bulk copy reads through the
// async proxy, and ordering a generic-proxy write before it needs fence.proxy.async in the WRITING thread, which waits for the
// thread's stores to reach L2: measured +0.34 µs per step, twice the step (profiles/seq_variants_r2.log).  With one proxy the
// ordering is ordinary release/acquire at CTA scope: st.global → warp barrier → mbarrier.arrive (release) on the "written"
// barrier of step k → the try_wait (acquire) of every lane of the TABLE PRODUCER warp on it before the lane issues its copies
// for step k + D.  Repeats at distance ≥ D are ordered that way, repeats at distance ≤ CIAO_HAZARD_WINDOW − 1 carry the HAZARD
// flag and are re-read by the thread that wrote them: no gap and no timing assumption.  The table producer is a second
// producer warp: folded into the lane that drives the TMA ring and re-arms the exchange barrier, the copies made that lane's
// iteration (831 cycles) longer than a step and it became the bottleneck (profiles/prof_seq_table_r2.log).  The rows a_i,
// their tails and the pass-written scalars are never written inside the kernel and stay on TMA.
static_assert(SEQ_D + 1 <= CIAO_HAZARD_WINDOW - 1, "every repeat closer than the ordered distance D + 1 must carry the HAZARD flag");
constexpr int SEQ_MAX_PART = 128;  // C·W ≤ 128
constexpr int SEQ_SLOT_EXTRA = 12;  // staged scalars [0,10) + index word [10] + pad (slots stay 16-byte aligned)
constexpr int SEQ_IDX_POS = 10;
// scalar area of a slot:  record tail at [0,6)  (b | λ | γ | γ/N | γ̂/γ | 0),  and/or the pass-written {b, λ, 0, c_i(z_full)}:
//   SVRG with cached c_i:    ss at [0,4)             → b, λ at 0, 1 as in the tail, c at 3        (2 bulk copies per step)
//   LFinito with cached c_i: tail at [0,6), ss at [6,10) → γ̂/γ at 4, c at 9                        (3 bulk copies per step)


// CZ: c_i(z_full) is read from the record tail (written by the last full-gradient pass at z_full)
// instead of being recomputed from a second dot product a_i·z_full.
template <int CPT, int ALG, int LOSS, int REG, bool CZ>
__global__ void __launch_bounds__(320, 1) seq_kernel(const SeqArgs p) {
    constexpr bool TABLE = (ALG == ALG_SAGA || ALG == ALG_FINITO);
    constexpr bool USES_ZFULL = (ALG == ALG_SVRG || ALG == ALG_LFINITO);
    constexpr bool TWO_DOTS = USES_ZFULL && !CZ;
    constexpr int D = SEQ_D;
    constexpr int H = CPT / 2;

    // an out-of-range index (flagged by prep_indices_kernel, same stream) must not touch z, av or the table: the reference throws
    // BoundsError before any state changes (SAGA_basic.jl:56).  Uniform over the cluster, so nobody is left at a barrier.
    if (*reinterpret_cast<const volatile int *>(p.err) != 0) return;
    // Cross-SM DSMEM latency is a per-SM-pair constant that differs between SM sets and between GPUs (seq_floor.cu), and a step
    // ends with the slowest pair of the cluster.  The kernel is therefore launched as a full grid of clusters — CTA → SM
    // placement of an idle GPU is deterministic — of which only the one at the calibrated position works.
    if ((int)(blockIdx.x / cluster_nctarank()) != p.active_cluster) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const bool TT = TABLE && p.table_tma;             // table row slices are staged in the ring slots, behind the record tail
    const int Tc = blockDim.x - (TT ? 64 : 32);       // compute threads; then the producer warp and (TT) the table producer warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = Tc >> 5;
    const uint32_t rank = cluster_ctarank(), C = cluster_nctarank();
    const int64_t dc = p.dc;
    const int cover = Tc * CPT;                       // columns covered by the thread grid (≥ dc)
    const size_t slot_doubles = (size_t)cover + SEQ_SLOT_EXTRA + (TT ? cover : 0);
    double *ring = reinterpret_cast<double *>(smem_raw);
    double *part = ring + D * slot_doubles;           // [2][npart_pad][2]
    uint64_t *row_bar = reinterpret_cast<uint64_t *>(part + 2 * (size_t)p.npart_pad * 2);
    uint64_t *part_bar = row_bar + D;
    uint64_t *wr_bar = part_bar + 2;                  // [D]: "table row of step k written" (W arrivals, TABLE kernels)
    const int64_t K = p.K;
    const int64_t cbase = (int64_t)rank * dc;
    const uint32_t part_bytes = C * W * 16;

    // zero the ring padding and the unused partial slots once
    for (size_t i = tid; i < D * slot_doubles + 2 * (size_t)p.npart_pad * 2; i += blockDim.x) ring[i] = 0.0;
    __shared__ long long clk_start[2];   // in shared memory, not registers: the step loop's register allocation must not change (§6 of the history)
    if (tid == 0) {
        uint32_t sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        p.smid_out[rank] = (int)sm;
        uint64_t ns0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
        clk_start[0] = clock64();
        clk_start[1] = (long long)ns0;
        for (int s = 0; s < D; ++s) mbar_init(&row_bar[s], TT ? 33 : 1);  // + one cp.async arrive per producer lane
        mbar_init(&part_bar[0], 1);  // one local arrive (expect_tx) per phase; the data arrives as tx bytes
        mbar_init(&part_bar[1], 1);
        if (TABLE)
            for (int s = 0; s < D; ++s) mbar_init(&wr_bar[s], W);
        fence_mbar_init();
    }
    fence_proxy_async();  // the zero fill (generic proxy) is ordered before the TMA writes into the ring
    __syncthreads();
    cluster_sync_all();   // peers' barriers exist before anyone sends to them

    if (warp == W) {
        // ===================== producer warp =====================
        // One lane drives the ring.  Its loop has to turn around faster than the compute warps' step (it re-arms the
        // exchange barrier and refills one slot per step), so it is kept lean: 32-bit counters, shared-memory addresses
        // and byte counts computed once, no shard search when all rows are local, two bulk copies per step for the
        // table-free algorithms.  (Profile: a 64-bit/three-copy version took ≈ 650 cycles per iteration and was the
        // bottleneck of every variant whose compute path is shorter than that.)
        if (lane == 0) {
            const int Ki = (int)K;  // run_seq_alg guarantees K < 2^31
            const uint32_t ring_s = smem_u32(ring), row_bar_s = smem_u32(row_bar), part_bar_s = smem_u32(part_bar);
            const uint32_t slot_bytes = (uint32_t)(slot_doubles * 8);
            const uint32_t row_bytes = (uint32_t)(dc * 8);
            constexpr bool SS_ONLY = CZ && ALG == ALG_SVRG;     // the pass-written scalars replace the record tail
            constexpr bool SS_EXTRA = CZ && ALG == ALG_LFINITO;  // … or come next to it
            const uint32_t tail_off = (uint32_t)cover * 8, idx_off = (uint32_t)(cover + SEQ_IDX_POS) * 8;
            const uint32_t tail_bytes = (uint32_t)(CIAO_TAIL_USED * 8);
            const uint32_t tx_bytes = row_bytes + (SS_ONLY ? 32u : tail_bytes + (SS_EXTRA ? 32u : 0u));
            const bool one_shard = p.rows.n <= 1;
            const double *base0 = p.rows.base[0] + cbase;
            const double *table0 = TABLE ? p.table + cbase : nullptr;
            const int64_t ld = p.ld, d_pad = p.d_pad;
            auto issue_row = [&](int step, int64_t pidx) {
                const int64_t i = pidx & CIAO_IDX_MASK;
                const uint32_t slot = (uint32_t)step & (D - 1);
                const uint32_t dst = ring_s + slot * slot_bytes, bar = row_bar_s + slot * 8;
                const double *src;
                if (one_shard) {
                    src = base0 + i * ld;
                } else {  // shard holding row i (≤ 8 shards: linear search)
                    int sh = 0;
                    while (sh + 1 < p.rows.n && i >= p.rows.start[sh + 1]) ++sh;
                    src = p.rows.base[sh] + (i - p.rows.start[sh]) * ld + cbase;
                }
                sts_b64(dst + idx_off, pidx);  // released by the arrive below
                mbar_arrive_expect_tx_s(bar, tx_bytes);
                tma_load_1d_s(dst, src, row_bytes, bar);
                // the step's scalars: the record tail and/or {b_i, λ_i, 0, c_i(z_full)} from the dense array of the last pass
                if (!SS_ONLY) tma_load_1d_s(dst + tail_off, src + (d_pad - cbase), tail_bytes, bar);
                if (SS_ONLY) tma_load_1d_s(dst + tail_off, p.ss + 4 * i, 32, bar);
                if (SS_EXTRA) tma_load_1d_s(dst + tail_off + CIAO_TAIL_USED * 8, p.ss + 4 * i, 32, bar);
                if (TABLE && !TT) tma_prefetch_l2(table0 + i * d_pad, row_bytes);
            };
            if (Ki > 0) mbar_arrive_expect_tx_s(part_bar_s, part_bytes);
            if (Ki > 1) mbar_arrive_expect_tx_s(part_bar_s + 8, part_bytes);
            for (int st = 0; st < D && st < Ki; ++st) issue_row(st, __ldg(p.idx + st));
            int64_t n1 = (D < Ki) ? __ldg(p.idx + D) : 0, n2 = (D + 1 < Ki) ? __ldg(p.idx + D + 1) : 0;
            const int64_t *idx_ahead = p.idx + D + 2;
#ifdef CIAO_SEQ_PROFILE
            long long prod_busy = 0;
#endif
            for (int k = 0; k < Ki; ++k) {
                const uint32_t pb = part_bar_s + ((uint32_t)k & 1u) * 8;
                // phase k complete ⇒ every warp of the cluster has consumed row k and sent its partial
                mbar_wait_s(pb, ((uint32_t)k >> 1) & 1u);
#ifdef CIAO_SEQ_PROFILE
                const long long pw = clock64();
#endif
                if (k + 2 < Ki) mbar_arrive_expect_tx_s(pb, part_bytes);  // arm the exchange of step k+2
                if (k + D < Ki) issue_row(k + D, n1);
                n1 = n2;
                n2 = (k + D + 2 < Ki) ? __ldg(idx_ahead + k) : 0;
#ifdef CIAO_SEQ_PROFILE
                prod_busy += clock64() - pw;
#endif
            }
#ifdef CIAO_SEQ_PROFILE
            g_seq_prof[rank * 8 + 5] = prod_busy;  // producer: cycles from wake-up to the end of its iteration
#endif
        }
    } else if (TABLE && warp == W + 1) {
        // ===================== table producer warp (TT kernels only) =====================
        // (`TABLE &&`: compiled out of the table-free kernels.  Present but never taken, this branch changed ptxas' register
        // allocation of the SVRG step loop — 113 instead of 116 registers — and cost 10 %: 0.311 → 0.343 µs/step at d = 4096,
        // profiles/seq_kernel_history_r2.md.)
        // Every lane copies 16-byte chunks of the table row slice of step k + D with cp.async (generic proxy) once the
        // "written" barrier of step k is complete: every compute warp has then finished step k, i.e. consumed the slot's
        // previous contents (row k) and released its table writes of steps ≤ k.
        const int Ki = (int)K;
        const uint32_t ring_s = smem_u32(ring), row_bar_s = smem_u32(row_bar), wr_bar_s = smem_u32(wr_bar);
        const uint32_t slot_bytes = (uint32_t)(slot_doubles * 8);
        const uint32_t table_off = (uint32_t)(cover + SEQ_SLOT_EXTRA) * 8;
        const uint32_t n_chunks = (uint32_t)(dc / 2);
        const double *table0 = p.table + cbase;
        const int64_t d_pad = p.d_pad;
        auto issue_table = [&](int step, int64_t pidx) {
            const uint32_t slot = (uint32_t)step & (D - 1);
            const uint32_t dst = ring_s + slot * slot_bytes + table_off, bar = row_bar_s + slot * 8;
            const double *trow = table0 + (pidx & CIAO_IDX_MASK) * d_pad;
            for (uint32_t ch = lane; ch < n_chunks; ch += 32) cp_async_16(dst + ch * 16, trow + 2 * ch);
            cp_async_arrive_noinc(bar);
        };
        for (int st = 0; st < D && st < Ki; ++st) issue_table(st, __ldg(p.idx + st));
        int64_t n1 = (D < Ki) ? __ldg(p.idx + D) : 0, n2 = (D + 1 < Ki) ? __ldg(p.idx + D + 1) : 0;
        const int64_t *idx_ahead = p.idx + D + 2;
        for (int k = 0; k + D < Ki; ++k) {
            mbar_wait_s(wr_bar_s + ((uint32_t)k & (D - 1)) * 8, ((uint32_t)k / D) & 1u);   // acquire, every lane
            issue_table(k + D, n1);
            n1 = n2;
            n2 = (k + D + 2 < Ki) ? __ldg(idx_ahead + k) : 0;
        }
    } else {
        // ===================== compute warps =====================
        int lcol[H];
        int64_t gcol[H];
        bool valid[H];
        double z[CPT], av[CPT], zf[CPT], zs[CPT], blo[CPT], bhi[CPT];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            lcol[h] = 2 * (tid + Tc * h);
            valid[h] = lcol[h] < dc;
            gcol[h] = cbase + (valid[h] ? lcol[h] : 0);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int q = 2 * h + e;
                const bool v = valid[h];
                const int64_t g = gcol[h] + e;
                z[q] = v ? (ALG == ALG_SVRG ? p.v_w[g] : p.v_z[g]) : 0.0;  // the running iterate: SVRG → w, others → z
                av[q] = v ? p.v_av[g] : 0.0;
                zf[q] = (v && USES_ZFULL) ? p.v_zfull[g] : 0.0;
                zs[q] = (v && ALG == ALG_SVRG) ? p.v_zsum[g] : 0.0;
                blo[q] = (v && REG == CIAO_REG_INDBOX && p.reg.lo_v) ? p.reg.lo_v[g] : p.reg.lo_s;
                bhi[q] = (v && REG == CIAO_REG_INDBOX && p.reg.hi_v) ? p.reg.hi_v[g] : p.reg.hi_s;
            }
        }
        const double gstep = (ALG == ALG_SVRG || ALG == ALG_SAGA) ? p.gamma : p.hat_gamma;
        const double gl = gstep * p.reg.lambda;
        const double cN = __ddiv_rn(p.hat_gamma, p.Nd);  // LFinito: γ̂/N
        const double rN = __ddiv_rn(1.0, p.Nd);
        const int E = p.npart_pad >> 5;                  // partials per lane in the final sum

        struct RowRegs {
            double a[CPT];
            double b, lam, gn, hg, cz;  // tail: b_i | λ_i | γ_i/N | γ̂/γ_i | c_i(z_full)
            int64_t ik;                 // prepared index word (row | flags)
            double2 t[H];               // this thread's slice of the table row s_i (SAGA/Finito)
        };
        // waits for the staged row of `step` and pulls this thread's slice + the scalars into registers
        auto load_row = [&](int64_t step, RowRegs &r) {
            const int slot = (int)(step & (D - 1));
            mbar_wait(&row_bar[slot], (uint32_t)((step / D) & 1));
            const double *rp = ring + slot * slot_doubles;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const double2 v = *reinterpret_cast<const double2 *>(rp + lcol[h]);
                r.a[2 * h] = v.x;
                r.a[2 * h + 1] = v.y;
            }
            r.b = rp[cover + TAIL_B];
            r.lam = rp[cover + TAIL_LAM];
            r.gn = (ALG == ALG_FINITO) ? rp[cover + TAIL_GAM_N] : 0.0;
            r.hg = (ALG == ALG_FINITO || ALG == ALG_LFINITO) ? rp[cover + TAIL_HAT_GAM] : 0.0;
            r.cz = (USES_ZFULL && CZ) ? rp[cover + (ALG == ALG_SVRG ? 3 : CIAO_TAIL_USED + 3)] : 0.0;
            r.ik = (ALG != ALG_SVRG) ? *reinterpret_cast<const int64_t *>(rp + cover + SEQ_IDX_POS) : 0;
            if (TT) {  // the table row slice was staged D steps ago (stale if the row was rewritten since: HAZARD flag)
#pragma unroll
                for (int h = 0; h < H; ++h) r.t[h] = *reinterpret_cast<const double2 *>(rp + cover + SEQ_SLOT_EXTRA + lcol[h]);
            }
        };
        // Fallback when the ring with table slices does not fit in shared memory (!TT): the table row of the next step is
        // pulled into the ping-pong register set with ld.global one step ahead (L2 prefetch issued D steps ago by the
        // producer).  Slower: ptxas tracks both register sets' loads with one scoreboard, so the update of step k also waits
        // for the load issued for step k+1 — the reason the TMA-staged ring is the default.
        auto load_table = [&](RowRegs &r) {
            const double *trow = p.table + (r.ik & CIAO_IDX_MASK) * p.d_pad;
#pragma unroll
            for (int h = 0; h < H; ++h)
                r.t[h] = valid[h] ? __ldcg(reinterpret_cast<const double2 *>(trow + gcol[h])) : make_double2(0.0, 0.0);
        };

        // this lane's remote destinations in the exchange (lane l < C talks to CTA l): computed once, not per step
        uint32_t send_dst[2], send_bar[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t peer = lane < C ? lane : 0;
            send_dst[q] = mapa_u32(smem_u32(part + ((size_t)q * p.npart_pad + rank * W + warp) * 2), peer);
            send_bar[q] = mapa_u32(smem_u32(&part_bar[q]), peer);
        }
        RowRegs rowA, rowB;
        if (K > 0) {
            load_row(0, rowA);
            if (TABLE && !TT) load_table(rowA);
        }
#ifdef CIAO_SEQ_PROFILE
        long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
        // one step; cur = registers of step k, nxt = registers to fill for step k+1
        auto step = [&](const int64_t k, RowRegs &cur, RowRegs &nxt) {
            const int par = (int)(k & 1);
            PROF_T(t_a);
            if (ALG == ALG_LFINITO && (cur.ik & CIAO_FLAG_PROX)) {  // Finito_LFinito.jl:92
#pragma unroll
                for (int q = 0; q < CPT; ++q) z[q] = prox_elem<REG>(av[q], gl, blo[q], bhi[q]);
            }
            // ---- dots: v0 = a·(w|z), v1 = a·z_full ------------------------------------
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int q = 0; q < CPT; ++q) {
                v0 = fma(cur.a[q], z[q], v0);
                if (TWO_DOTS) v1 = fma(cur.a[q], zf[q], v1);
            }
            v0 = warp_sum_mma(v0, lane);
            if (TWO_DOTS) v1 = warp_sum_mma(v1, lane);
            PROF_T(t_b);
            if (lane < C) st_async_v2f64(send_dst[par], v0, v1, send_bar[par]);
            // ---- in the shadow of the exchange: ∇f_i(z_full) = c_i(z_full)·a_i from the cached scalar (the step is fp64-issue
            //      bound after the exchange, so every product formed here comes off the critical path)
            // ptxas sinks register-only arithmetic below the wait loop; making the wait's parity operand depend on the products
            // (through a mask that is zero at run time) keeps them here
            double gz[CPT];
            int pin = 0;
            if (CZ) {
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
                    gz[q] = grad_elem<LOSS>(cur.a[q], cur.cz, cur.lam);
                    if (ALG == ALG_LFINITO) gz[q] = __dmul_rn(cN, gz[q]);
                    pin ^= __double2hiint(gz[q]);
                }
            }
            // ---- while the exchange is in flight: the next row and its table row go to registers
            if (k + 1 < K) {
#ifdef CIAO_SEQ_PROFILE
                {
                    const long long w0 = clock64();
                    mbar_wait(&row_bar[(k + 1) & (D - 1)], (uint32_t)(((k + 1) / D) & 1));
                    prof_acc[4] += clock64() - w0;  // wait for the staged row of the next step (included in phase 1)
                }
#endif
                load_row(k + 1, nxt);
                if (TABLE && !TT) load_table(nxt);
            }
            PROF_T(t_c);
#ifdef CIAO_SEQ_SPIN
            {   // experiment: poll the phase with the non-blocking test instead of the suspending try_wait
                const uint32_t ph = (uint32_t)((k >> 1) & 1) ^ (uint32_t)(pin & p.zero);
                while (!mbar_test(&part_bar[par], ph)) {}
            }
#else
            mbar_wait(&part_bar[par], (uint32_t)((k >> 1) & 1) ^ (uint32_t)(pin & p.zero));
#endif
            PROF_T(t_d);
            // every warp reduces the same C·W partials with the same tensor-core sum → identical bits everywhere
            double u0 = 0.0, u1 = 0.0;
            {
                const double2 *pp = reinterpret_cast<const double2 *>(part + (size_t)par * p.npart_pad * 2) + lane;
                const double2 v = pp[0];  // E == 1 (C·W ≤ 32) is the common shape: no loop, no branches
                u0 = v.x;
                if (TWO_DOTS) u1 = v.y;
#ifndef CIAO_SEQ_E1   // experiment CIAO_SEQ_E1: C·W ≤ 32 assumed at compile time
                for (int e = 1; e < E; ++e) {
                    const double2 w = pp[e * 32];
                    u0 += w.x;
                    if (TWO_DOTS) u1 += w.y;
                }
#endif
                u0 = warp_sum_mma(u0, lane);
                if (TWO_DOTS) u1 = warp_sum_mma(u1, lane);
            }
            const double tb = cur.b, tl = cur.lam;

            // ---- fused update ---------------------------------------------------------
            if (ALG == ALG_SVRG) {  // SVRG_basic.jl:74-81
                const double cz = CZ ? cur.cz : loss_coef<LOSS>(u1, tb, tl);
                const double cw = loss_coef<LOSS>(u0, tb, tl);
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
                    double t = __dsub_rn(CZ ? gz[q] : grad_elem<LOSS>(cur.a[q], cz, tl), grad_elem<LOSS>(cur.a[q], cw, tl));
                    t = __dsub_rn(t, av[q]);
                    t = __dmul_rn(t, p.gamma);
                    t = __dadd_rn(t, z[q]);
                    z[q] = prox_elem<REG>(t, gl, blo[q], bhi[q]);
                    zs[q] = __dadd_rn(zs[q], z[q]);
                }
            } else if (ALG == ALG_LFINITO) {  // Finito_LFinito.jl:94-98
                const double czf = CZ ? cur.cz : loss_coef<LOSS>(u1, tb, tl);
                const double czz = loss_coef<LOSS>(u0, tb, tl);
                const double rr = cur.hg;  // γ̂/γ_i
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
                    av[q] = __dadd_rn(av[q], CZ ? gz[q] : __dmul_rn(cN, grad_elem<LOSS>(cur.a[q], czf, tl)));
                    av[q] = __dsub_rn(av[q], __dmul_rn(cN, grad_elem<LOSS>(cur.a[q], czz, tl)));
                    av[q] = __dadd_rn(av[q], __dmul_rn(rr, __dsub_rn(z[q], zf[q])));
                }
            } else {
                const double c = loss_coef<LOSS>(u0, tb, tl);
                double *trow = p.table + (cur.ik & CIAO_IDX_MASK) * p.d_pad;
                if (cur.ik & CIAO_FLAG_HAZARD) {  // the row was rewritten after its prefetch was issued
#pragma unroll
                    for (int h = 0; h < H; ++h)
                        if (valid[h]) cur.t[h] = __ldcg(reinterpret_cast<const double2 *>(trow + gcol[h]));
                }
                double snew[CPT];
                if (ALG == ALG_SAGA) {  // SAGA_basic.jl:56-65
                    // the SAG/SAGA choice is made once per step, outside the element loop: a branch per element keeps the
                    // compiler from interleaving the four independent element chains (measured: 520 → 350 cycles at d = 4096)
                    if (p.sag) {
#pragma unroll
                        for (int q = 0; q < CPT; ++q) {
                            const double so = (q & 1) ? cur.t[q / 2].y : cur.t[q / 2].x;
                            const double g = grad_elem<LOSS>(cur.a[q], c, tl);
                            av[q] = __dadd_rn(av[q], div_by(__dsub_rn(g, so), p.Nd, rN));      // :58
                            const double w = __dsub_rn(z[q], __dmul_rn(p.gamma, av[q]));       // :59
                            z[q] = prox_elem<REG>(w, gl, blo[q], bhi[q]);
                            snew[q] = g;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < CPT; ++q) {
                            const double so = (q & 1) ? cur.t[q / 2].y : cur.t[q / 2].x;
                            const double g = grad_elem<LOSS>(cur.a[q], c, tl);
                            const double diff = __dsub_rn(g, so);
                            const double w = __dsub_rn(z[q], __dmul_rn(p.gamma, __dadd_rn(diff, av[q])));  // :61
                            av[q] = __dadd_rn(av[q], div_by(diff, p.Nd, rN));                  // :62
                            z[q] = prox_elem<REG>(w, gl, blo[q], bhi[q]);
                            snew[q] = g;
                        }
                    }
                } else {  // Finito_basic.jl:112-118
                    const double cneg = -cur.gn;  // −(γ_i/N)
                    const double rr = cur.hg;     // γ̂/γ_i
#pragma unroll
                    for (int q = 0; q < CPT; ++q) {
                        const double so = (q & 1) ? cur.t[q / 2].y : cur.t[q / 2].x;
                        double t = __dmul_rn(grad_elem<LOSS>(cur.a[q], c, tl), cneg);
                        t = __dadd_rn(t, z[q]);
                        av[q] = __dadd_rn(av[q], __dmul_rn(__dsub_rn(t, so), rr));
                        snew[q] = t;
                    }
                    if (cur.ik & CIAO_FLAG_PROX) {
#pragma unroll
                        for (int q = 0; q < CPT; ++q) z[q] = prox_elem<REG>(av[q], gl, blo[q], bhi[q]);
                    }
                }
#pragma unroll
                for (int h = 0; h < H; ++h)
                    if (valid[h])
                        __stcg(reinterpret_cast<double2 *>(trow + gcol[h]), make_double2(snew[2 * h], snew[2 * h + 1]));
                if (TT) {   // release: this warp's table writes of step k are visible to whoever acquires wr_bar
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&wr_bar[k & (D - 1)]);
                }
            }
            PROF_T(t_e);
            PROF_ADD(0, t_a, t_b);  // dots + warp shuffles
            PROF_ADD(1, t_b, t_c);  // send + products pinned in the shadow + registers of the next step
            PROF_ADD(2, t_c, t_d);  // remaining wait for the cluster exchange
            PROF_ADD(3, t_d, t_e);  // sum of the partials + fused update
        };
        int64_t k = 0;
        for (; k + 1 < K; k += 2) {  // ping-pong the row registers: no copies between steps
            step(k, rowA, rowB);
            step(k + 1, rowB, rowA);
        }
        if (k < K) step(k, rowA, rowB);

        // ---- epilogue: state back to HBM ------------------------------------------------
#pragma unroll
        for (int h = 0; h < H; ++h) {
            if (!valid[h]) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int q = 2 * h + e;
                const int64_t g = gcol[h] + e;
                if (ALG == ALG_SVRG) {  // SVRG_basic.jl:84-86
                    const double zfull = __ddiv_rn(zs[q], p.m_d);
                    p.v_zfull[g] = zfull;
                    p.v_w[g] = p.plus ? z[q] : zfull;
                    p.v_zsum[g] = 0.0;
                } else {
                    p.v_z[g] = z[q];
                    p.v_av[g] = av[q];
                }
            }
        }
#ifdef CIAO_SEQ_PROFILE
        if (tid == 0)
            for (int i = 0; i < 5; ++i) g_seq_prof[rank * 8 + i] = prof_acc[i];  // [5] belongs to the producer lane
#endif
    }
    cluster_sync_all();  // nobody exits while a peer may still touch its shared memory
    if (tid == 0 && rank == 0) {
        uint64_t ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        p.clk_out[0] = clock64() - clk_start[0];
        p.clk_out[1] = (long long)ns1 - clk_start[1];
    }
}

// ---------------------------------------------------------------------------
struct SeqShape {
    int C, Tc, cpt, npart_pad;
    int64_t dc;
};

static size_t seq_smem_bytes(const SeqShape &sh, bool table_in_ring) {
    const size_t cover = (size_t)sh.Tc * sh.cpt;
    return (size_t)SEQ_D * (cover + SEQ_SLOT_EXTRA + (table_in_ring ? cover : 0)) * 8 + 2 * (size_t)sh.npart_pad * 2 * 8 + (2 * SEQ_D + 2) * 8 + 128;
}

template <int CPT, int ALG, int LOSS, int REG, bool CZ>
static int launch_seq(ciao_ctx *c, const SeqArgs &a, const SeqShape &sh) {
    auto kern = seq_kernel<CPT, ALG, LOSS, REG, CZ>;
    const size_t smem = seq_smem_bytes(sh, a.table_tma != 0);
    static size_t configured[CIAO_MAX_DEVICES] = {};
    if (smem > configured[c->device % CIAO_MAX_DEVICES]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[c->device % CIAO_MAX_DEVICES] = smem;
    }
    static bool big_cluster[CIAO_MAX_DEVICES] = {};
    if (sh.C > 8 && !big_cluster[c->device % CIAO_MAX_DEVICES]) {  // 16 CTAs: non-portable cluster size (opt-in on sm_100)
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        big_cluster[c->device % CIAO_MAX_DEVICES] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sh.C * (a.active_cluster + 1));   // clusters 0 … active − 1 exit at once
    cfg.blockDim = dim3(sh.Tc + (a.table_tma ? 64 : 32));   // + producer warp (+ table producer warp)
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = sh.C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    return CIAO_OK;
}

template <int CPT, int ALG, int LOSS, int REG>
static int launch_seq_cz(ciao_ctx *c, const SeqArgs &a, const SeqShape &sh) {
    if constexpr (ALG == ALG_SVRG || ALG == ALG_LFINITO) {
        if (c->cz_valid) return launch_seq<CPT, ALG, LOSS, REG, true>(c, a, sh);
    }
    return launch_seq<CPT, ALG, LOSS, REG, false>(c, a, sh);
}

template <int CPT, int ALG, int LOSS>
static int launch_seq_reg(ciao_ctx *c, const SeqArgs &a, const SeqShape &sh) {
    switch (c->reg.kind) {
        case CIAO_REG_NORML1: return launch_seq_cz<CPT, ALG, LOSS, CIAO_REG_NORML1>(c, a, sh);
        case CIAO_REG_INDBOX: return launch_seq_cz<CPT, ALG, LOSS, CIAO_REG_INDBOX>(c, a, sh);
        default: return launch_seq_cz<CPT, ALG, LOSS, CIAO_REG_ZERO>(c, a, sh);
    }
}

template <int CPT, int ALG>
static int launch_seq_loss(ciao_ctx *c, const SeqArgs &a, const SeqShape &sh) {
    return c->loss_kind == CIAO_LOSS_LS ? launch_seq_reg<CPT, ALG, CIAO_LOSS_LS>(c, a, sh)
                                        : launch_seq_reg<CPT, ALG, CIAO_LOSS_LOGISTIC>(c, a, sh);
}

// Chooses the cluster shape: C CTAs × Tc compute threads × CPT columns per thread cover d_pad.
static int seq_shape(ciao_ctx *c, SeqShape *sh) {
    const int64_t d_pad = c->d_pad;
    int C = c->seq_cluster > 0 ? c->seq_cluster : (d_pad >= 1024 ? 8 : (d_pad >= 256 ? 4 : 1));  // measured: profiles/tune_seq_r1.json
    while (C > 1 && (d_pad % (4 * C) != 0)) C >>= 1;
    const int64_t dc = d_pad / C;
    const int T_target = c->seq_threads > 0 ? std::min(c->seq_threads, 256) : 128;
    int cpt = 2;
    while (cpt < 8 && (dc + cpt - 1) / cpt > T_target) cpt *= 2;
    const int64_t T = ((dc + cpt - 1) / cpt + 31) / 32 * 32;
    if (T > 256 || C * (T / 32) > SEQ_MAX_PART)
        CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "sequential kernel: d = %lld too large for cluster %d", (long long)c->d, C);
    sh->C = C; sh->Tc = (int)T; sh->cpt = cpt; sh->dc = dc;
    sh->npart_pad = (C * (int)(T / 32) + 31) / 32 * 32;
    return CIAO_OK;
}

// one translation unit per algorithm (seq_svrg.cu, …) instantiates this
template <int ALG>
static int run_seq_alg(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, double m_d) {
    NvtxRange nvtx(ALG == ALG_SVRG ? "ciao:seq:svrg" : ALG == ALG_SAGA ? "ciao:seq:saga" : ALG == ALG_FINITO ? "ciao:seq:finito" : "ciao:seq:lfinito");
    if (K <= 0) return CIAO_OK;
    if (K >= (int64_t)1 << 31) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "sequential kernel: more than 2^31 - 1 steps in one call");
    SeqShape sh;
    CIAO_TRY(seq_shape(c, &sh));
    SeqArgs a;
    if (c->peers.n > 1) {
        a.rows = c->peers;
    } else {
        a.rows.n = 1;
        a.rows.base[0] = c->rec;
        a.rows.start[0] = 0;
        a.rows.start[1] = c->n_rows;
    }
    a.ld = c->ld; a.d_pad = c->d_pad; a.dc = sh.dc;
    a.idx = idx_prepared; a.K = K; a.table = c->table; a.ss = c->ss;
    a.v_z = ctx_vec(c, CIAO_VEC_Z); a.v_zfull = ctx_vec(c, CIAO_VEC_Z_FULL); a.v_w = ctx_vec(c, CIAO_VEC_W);
    a.v_av = ctx_vec(c, CIAO_VEC_AV); a.v_zsum = ctx_vec(c, CIAO_VEC_Z);  // SVRG: state.z is the running sum of inner iterates
    a.gamma = c->gamma; a.hat_gamma = c->hat_gamma; a.Nd = (double)c->N_total; a.m_d = m_d;
    a.plus = c->plus; a.sag = c->sag; a.reg = c->reg; a.npart_pad = sh.npart_pad; a.zero = 0; a.err = c->err_dev;
    a.smid_out = c->seq_smid; c->seq_smid_n = sh.C;
    a.clk_out = reinterpret_cast<long long *>(c->seq_smid + 16);
    a.active_cluster = std::max(0, std::min(c->seq_cluster_pos, c->num_sms / sh.C - 1));
    a.table_tma = (ALG == ALG_SAGA || ALG == ALG_FINITO) && seq_smem_bytes(sh, true) <= 200 * 1024 && !c->seq_table_ldg;
    CUDA_TRY(cudaEventRecord(c->ev_sa, c->stream));
    int rc;
    switch (sh.cpt) {
        case 2: rc = launch_seq_loss<2, ALG>(c, a, sh); break;
        case 4: rc = launch_seq_loss<4, ALG>(c, a, sh); break;
        default: rc = launch_seq_loss<8, ALG>(c, a, sh); break;
    }
    CIAO_TRY(rc);
    CUDA_TRY(cudaEventRecord(c->ev_sb, c->stream));
    c->timing.launches += 1;
    c->timing.last_seq_steps = K;
    c->seq_timed = true;
    return CIAO_OK;
}

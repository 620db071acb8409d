// seq_adaptive.cu — adaptive Finito (Finito_adaptive.jl): the backtracking linesearch on γ_i and the table update as ONE
// persistent thread-block-cluster kernel per call, plus the table/stepsize initialisation pass.
//
//   adaptive_init_kernel   s_i = x0, ∇f_i(x0), f_i(x0), γ_i = α / (‖∇f_i(x0+1) − ∇f_i(x0)‖ / (√d · N))     :65-87
//   adaptive_kernel        K steps of :101-160 on a host-generated index sequence
//
// Same skeleton as seq_impl.cuh: the d columns are split over the CTAs of a cluster, a thread keeps its
// columns of z and av in registers for the whole call, a producer lane TMA-stages row a_i, the record tail, the table row
// x_i and the per-component scalars {γ_i, f_i(x_i), c_i(x_i)} eight steps ahead, and every reduction is an all-to-all of
// warp partials through DSMEM (st.async) followed by the same tensor-core sum in every warp — so all threads of the
// cluster hold bit-identical scalars and take the data-dependent branches of the linesearch together, with no barrier.
// Differences: a step needs three reductions per linesearch trial (a_i·z, ⟨∇f_i(x_i), z − x_i⟩, ‖z − x_i‖²) and an unknown
// number of trials, so exchanges are counted separately from steps, thread 0 of every CTA re-arms the exchange barrier, and
// ring slots are released through an explicit "empty" mbarrier.  For row models ∇f_i(x_i) = c_i·a_i (rank one), so the
// reference's N×d gradient table is kept as the scalar c_i: the element values fl(fl(a_k·c_i)·λ_i) are reproduced on the fly.
// A component that recurs within the prefetch window reads its scalars from a warp-private history of the last 32 steps
// (prep_indices_kernel supplies the distance) and re-reads its table row behind its own store.
#include "seq_impl.cuh"

struct AdArgs {
    const double *rec;  // [N][ld] row records
    int64_t ld, d_pad, dc;
    const int64_t *idx;  // prepared: 0-based row | HAZARD | distance
    int64_t K;
    double *table;       // [N][d_pad]  x_i            (state.s)
    double *ad;          // [8][N][4]   {γ_i, f_i(x_i), c_i(x_i), 0}   (state.γ, state.fi_x, state.∇f as a scalar), one private
                         //             copy per CTA of the cluster: a CTA stages (cp.async) only what it wrote itself
    int64_t N;
    double *scal;        // [0] γ̂ (in/out)
    int64_t *counters;   // [0] steps completed by this call, [1] linesearch reductions of γ
    double *v_z, *v_av;
    double Nd, alpha, tol_b;
    int npart_pad;
    int zero;            // always 0, opaque to the compiler (see seq_impl.cuh)
    RegParams reg;
};

constexpr int AD_D = 8;       // ring depth
constexpr int AD_HIST = 32;   // steps of per-component scalar history (≥ CIAO_HAZARD_WINDOW)
constexpr int AD_EXTRA = 12;  // scalar area of a slot: record tail [0,6) | {γ_i, f_i, c_i, 0} [6,10) | index word [10] | pad
static_assert((AD_D & (AD_D - 1)) == 0 && AD_D + 2 <= CIAO_HAZARD_WINDOW - 1 && CIAO_HAZARD_WINDOW <= AD_HIST,
              "hazard window vs ring depth / history");

template <int CPT, int LOSS, int REG>
__global__ void __launch_bounds__(288, 1) adaptive_kernel(const AdArgs p) {
    constexpr int D = AD_D, H = CPT / 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Tc = blockDim.x - 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = Tc >> 5;
    const uint32_t rank = cluster_ctarank(), C = cluster_nctarank();
    const int64_t dc = p.dc, K = p.K;
    const int cover = Tc * CPT;
    const size_t slot_doubles = 2 * (size_t)cover + AD_EXTRA;
    double *ring = reinterpret_cast<double *>(smem_raw);
    double *part = ring + D * slot_doubles;                       // [2][npart_pad][4]
    double *hist = part + 2 * (size_t)p.npart_pad * 4;            // [W][AD_HIST][4]
    uint64_t *row_bar = reinterpret_cast<uint64_t *>(hist + (size_t)W * AD_HIST * 4);
    uint64_t *empty_bar = row_bar + D;
    uint64_t *part_bar = empty_bar + D;
    const int64_t cbase = (int64_t)rank * dc;
    const uint32_t part_bytes = C * W * 32;

    for (size_t i = tid; i < D * slot_doubles + 2 * (size_t)p.npart_pad * 4 + (size_t)W * AD_HIST * 4; i += blockDim.x) ring[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < D; ++s) {
            mbar_init(&row_bar[s], 33);   // lane 0's arrive.expect_tx + one cp.async arrive per producer lane
            mbar_init(&empty_bar[s], W);
        }
        mbar_init(&part_bar[0], 1);
        mbar_init(&part_bar[1], 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&part_bar[0], part_bytes);  // exchanges 0 and 1; exchange e+2 is armed when e completes
        mbar_arrive_expect_tx(&part_bar[1], part_bytes);
    }
    fence_proxy_async();
    __syncthreads();
    cluster_sync_all();

    if (warp == W) {
        // ===================== producer warp: keeps the ring D steps ahead =====================
        // Lane 0 waits for the slot and issues the bulk copies of the row and its tail (never written in the kernel); the table
        // row x_i and the scalars {γ_i, f_i, c_i} — both rewritten by the compute warps of THIS CTA — are copied by all lanes
        // with cp.async (generic proxy, see seq_impl.cuh).  Ordering: st.global of step k → warp barrier + release arrive on
        // empty_bar in the next load() → lane 0's acquire wait below → this warp's barrier → the copies.  Repeats at distance
        // ≥ AD_D + 2 are ordered that way, closer ones carry the HAZARD flag (history / re-read behind the own store).
        const int Ki = (int)K;
        const uint32_t ring_s = smem_u32(ring), row_bar_s = smem_u32(row_bar), empty_bar_s = smem_u32(empty_bar);
        const uint32_t slot_bytes = (uint32_t)(slot_doubles * 8), row_bytes = (uint32_t)(dc * 8);
        const uint32_t tail_off = (uint32_t)cover * 8, table_off = (uint32_t)(cover + AD_EXTRA) * 8;
        const uint32_t tx_bytes = row_bytes + CIAO_TAIL_USED * 8;
        const uint32_t n_chunks = (uint32_t)(dc / 2);
        const double *ad_mine = p.ad + (size_t)rank * 4 * (size_t)p.N;
        for (int st = 0; st < Ki; ++st) {
            const int64_t pidx = __ldg(p.idx + st);
            const int64_t i = pidx & CIAO_IDX_MASK;
            const uint32_t slot = (uint32_t)st & (D - 1);
            const uint32_t dst = ring_s + slot * slot_bytes, bar = row_bar_s + slot * 8;
            if (lane == 0) {
                // every compute warp has pulled step st − D out of this slot
                if (st >= D) mbar_wait_s(empty_bar_s + slot * 8, (((uint32_t)st / D) - 1u) & 1u);
                const double *src = p.rec + i * p.ld;
                sts_b64(dst + tail_off + 80, pidx);
                mbar_arrive_expect_tx_s(bar, tx_bytes);
                tma_load_1d_s(dst, src + cbase, row_bytes, bar);
                tma_load_1d_s(dst + tail_off, src + p.d_pad, CIAO_TAIL_USED * 8, bar);
            }
            __syncwarp();
            const double *trow = p.table + i * p.d_pad + cbase;
            for (uint32_t ch = lane; ch < n_chunks; ch += 32) cp_async_16(dst + table_off + ch * 16, trow + 2 * ch);
            if (lane < 2) cp_async_16(dst + tail_off + CIAO_TAIL_USED * 8 + lane * 16, ad_mine + 4 * i + 2 * lane);
            cp_async_arrive_noinc(bar);
        }
    } else {
        // ===================== compute warps =====================
        int lcol[H];
        int64_t gcol[H];
        bool valid[H];
        double z[CPT], av[CPT], blo[CPT], bhi[CPT];
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
                z[q] = v ? p.v_z[g] : 0.0;
                av[q] = v ? p.v_av[g] : 0.0;
                blo[q] = (v && REG == CIAO_REG_INDBOX && p.reg.lo_v) ? p.reg.lo_v[g] : p.reg.lo_s;
                bhi[q] = (v && REG == CIAO_REG_INDBOX && p.reg.hi_v) ? p.reg.hi_v[g] : p.reg.hi_s;
            }
        }
        double hg = p.scal[0];
        double gl = __dmul_rn(hg, p.reg.lambda);
        const double thr = __ddiv_rn(p.tol_b, p.Nd);                           // :124 tol_b / N
        const double half_N_alpha = __dmul_rn(__dmul_rn(0.5, p.Nd), p.alpha);  // :131 0.5 * N * α
        const double rNd = __drcp_rn(p.Nd);
        const int E = p.npart_pad >> 5;
        double *hist_w = hist + (size_t)warp * AD_HIST * 4;
        uint32_t send_dst[2], send_bar[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t peer = lane < C ? lane : 0;
            send_dst[q] = mapa_u32(smem_u32(part + ((size_t)q * p.npart_pad + rank * W + warp) * 4), peer);
            send_bar[q] = mapa_u32(smem_u32(&part_bar[q]), peer);
        }
        int64_t done = 0, nbt = 0;
        uint32_t ex = 0;  // exchanges so far (identical in every thread of the cluster)
        bool stopped = false;

        struct Row {
            double a[CPT], s[CPT];
            double tb, tl, gam, fix, cold;
            int64_t ik;
        };
        // pulls the staged data of `step` into registers and hands the slot back to the producer
        auto load = [&](int64_t step, Row &r) {
            const int slot = (int)(step & (D - 1));
            mbar_wait(&row_bar[slot], (uint32_t)((step / D) & 1));
            const double *rp = ring + slot * slot_doubles;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const double2 va = *reinterpret_cast<const double2 *>(rp + lcol[h]);
                const double2 vs = *reinterpret_cast<const double2 *>(rp + cover + AD_EXTRA + lcol[h]);
                r.a[2 * h] = va.x; r.a[2 * h + 1] = va.y;
                r.s[2 * h] = vs.x; r.s[2 * h + 1] = vs.y;
            }
            r.tb = rp[cover + TAIL_B]; r.tl = rp[cover + TAIL_LAM];
            r.gam = rp[cover + CIAO_TAIL_USED]; r.fix = rp[cover + CIAO_TAIL_USED + 1]; r.cold = rp[cover + CIAO_TAIL_USED + 2];
            r.ik = *reinterpret_cast<const int64_t *>(rp + cover + 10);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
        };
        Row nxt;
        bool have_next = false;
        if (K > 0) {
            load(0, nxt);
            have_next = true;
        }
#ifdef CIAO_SEQ_PROFILE
        long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
        for (int64_t k = 0; k < K; ++k) {
            PROF_T(t_0);
#ifdef CIAO_SEQ_PROFILE
            long long t_last = t_0;
#endif
            if (!have_next) load(k, nxt);   // only after a stop (slots are still drained) or when the prefetch below was skipped
            have_next = false;
            if (stopped) continue;
            double a[CPT], s[CPT];
#pragma unroll
            for (int q = 0; q < CPT; ++q) { a[q] = nxt.a[q]; s[q] = nxt.s[q]; }
            const double tb = nxt.tb, tl = nxt.tl;
            double gam = nxt.gam, fix = nxt.fix, cold = nxt.cold;
            const int64_t ik = nxt.ik;

            double *trow = p.table + (ik & CIAO_IDX_MASK) * p.d_pad;
            if (ik & CIAO_FLAG_HAZARD) {
                // the component was updated after its copies were staged: table row behind our own store, scalars from the history
#pragma unroll
                for (int h = 0; h < H; ++h)
                    if (valid[h]) {
                        const double2 vs = __ldcg(reinterpret_cast<const double2 *>(trow + gcol[h]));
                        s[2 * h] = vs.x; s[2 * h + 1] = vs.y;
                    }
                const double *hp = hist_w + ((k - ((ik >> CIAO_HAZ_DIST_SHIFT) & 31)) & (AD_HIST - 1)) * 4;
                gam = hp[0]; fix = hp[1]; cold = hp[2];
            }
            double gold[CPT], res[CPT];
#pragma unroll
            for (int q = 0; q < CPT; ++q) {
                gold[q] = grad_elem<LOSS>(a[q], cold, tl);  // ∇f_i(x_i)[k], the values the reference's table holds
                res[q] = __dsub_rn(z[q], s[q]);             // :121
            }
            double u = 0.0, fi_z = 0.0, r_main = 0.0, cN_main = 0.0;
            for (;;) {                                      // :123-147
                if (gam < thr) {                            // :124-127  `return nothing`
                    stopped = true;
                    break;
                }
                PROF_T(t_a);
                PROF_ADD(0, t_last, t_a);   // everything before the partial sums (first trial: step setup; later: backtracking)
                double p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
                    p1 = fma(a[q], z[q], p1);
                    p2 = fma(gold[q], res[q], p2);
                    p3 = fma(res[q], res[q], p3);
                }
                p1 = warp_sum_mma(p1, lane);
                p2 = warp_sum_mma(p2, lane);
                p3 = warp_sum_mma(p3, lane);
                const int par = (int)(ex & 1);
                {   // selects, not array indexing by a run-time parity (that puts the addresses in local memory)
                    const uint32_t sd = par ? send_dst[1] : send_dst[0], sb = par ? send_bar[1] : send_bar[0];
                    if (lane < C) {
                        st_async_v2f64(sd, p1, p2, sb);
                        st_async_v2f64(sd + 16, p3, 0.0, sb);
                    }
                }
                PROF_T(t_b);
                // in the shadow of the exchange: the three divisions that depend only on γ_i and γ̂ (pinned in front of the wait
                // through the parity operand, as in seq_impl.cuh), and the registers of the next step
                // one reciprocal + exact-remainder corrections (div_by: the correctly rounded quotients) instead of three divisions
                const double rgam = __drcp_rn(gam);
                const double coef = div_by(half_N_alpha, gam, rgam);   // :131  0.5·N·α / γ_i
                const double r_hg = div_by(hg, gam, rgam);             // :149  γ̂ / γ_i
                const double cN = div_by(hg, p.Nd, rNd);               // :151  γ̂ / N
                const int pin = __double2hiint(coef) ^ __double2hiint(r_hg) ^ __double2hiint(cN);
                if (!have_next && k + 1 < K) {
                    load(k + 1, nxt);
                    have_next = true;
                }
                PROF_T(t_c);
                mbar_wait(&part_bar[par], ((ex >> 1) & 1) ^ (uint32_t)(pin & p.zero));
                PROF_T(t_d);
                if (tid == 0) mbar_arrive_expect_tx(&part_bar[par], part_bytes);  // arm exchange ex + 2
                double U = 0.0, Dg = 0.0, R2 = 0.0;
                {
                    const double *pp = part + ((size_t)par * p.npart_pad + lane) * 4;
                    for (int e = 0; e < E; ++e) {
                        const double2 v01 = *reinterpret_cast<const double2 *>(pp + (size_t)e * 32 * 4);
                        const double v2 = pp[(size_t)e * 32 * 4 + 2];
                        U += v01.x; Dg += v01.y; R2 += v2;
                    }
                    U = warp_sum_mma(U, lane);
                    Dg = warp_sum_mma(Dg, lane);
                    R2 = warp_sum_mma(R2, lane);
                }
                ++ex;
                PROF_ADD(1, t_a, t_b);  // partial sums + warp sums + send
                PROF_ADD(2, t_b, t_c);  // shadow work: divisions, next-step registers
                PROF_ADD(3, t_c, t_d);  // remaining exchange wait
#ifdef CIAO_SEQ_PROFILE
                t_last = t_d;
#endif
                fi_z = loss_value<LOSS>(U, tb, tl);                                              // :128
                const double nr = __dsqrt_rn(R2);
                const double fi_model = __dadd_rn(__dadd_rn(fix, Dg), __dmul_rn(coef, __dmul_rn(nr, nr)));  // :129-132
                const double tol = __dmul_rn(10 * 2.220446049250313e-16, __dadd_rn(1.0, fabs(fi_z)));  // :133
                if (fi_z <= __dadd_rn(fi_model, tol)) {                                          // :134
                    u = U;
                    r_main = r_hg;
                    cN_main = cN;
                    break;
                }
                const double gam_b = gam;                                                        // :136
                gam = __dmul_rn(gam, 0.8);                                                       // :137
                const double hg_new = __ddiv_rn(1.0, __dsub_rn(__dadd_rn(__ddiv_rn(1.0, hg), __ddiv_rn(1.0, gam)), __ddiv_rn(1.0, gam_b)));  // :142
                const double gl_new = __dmul_rn(hg_new, p.reg.lambda);
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
                    double t = __ddiv_rn(av[q], hg);                                             // :139
                    t = __dadd_rn(t, __ddiv_rn(s[q], gam));                                      // :140
                    t = __dsub_rn(t, __ddiv_rn(s[q], gam_b));                                    // :141
                    av[q] = __dmul_rn(t, hg_new);                                                // :143
                    z[q] = prox_elem<REG>(av[q], gl_new, blo[q], bhi[q]);                        // :144
                    res[q] = __dsub_rn(z[q], s[q]);                                              // :145
                }
                hg = hg_new;
                gl = gl_new;
                ++nbt;
            }
            if (stopped) continue;
            PROF_T(t_e);
            PROF_ADD(4, t_last, t_e);  // totals + warp sums + model test of the accepted trial
            // ---- main step :149-154 ----
            const double r = r_main, cN = cN_main;   // γ̂/γ_i and γ̂/N of the accepted trial
            const double cnew = loss_coef<LOSS>(u, tb, tl);
#pragma unroll
            for (int q = 0; q < CPT; ++q) {
                double t = __dadd_rn(av[q], __dmul_rn(r, __dsub_rn(z[q], s[q])));                // :149
                t = __dadd_rn(t, __dmul_rn(cN, gold[q]));                                        // :151
                t = __dsub_rn(t, __dmul_rn(cN, grad_elem<LOSS>(a[q], cnew, tl)));                // :152-153
                av[q] = t;
            }
#pragma unroll
            for (int h = 0; h < H; ++h)                                                          // :150  x_i = z
                if (valid[h]) __stcg(reinterpret_cast<double2 *>(trow + gcol[h]), make_double2(z[2 * h], z[2 * h + 1]));
#pragma unroll
            for (int q = 0; q < CPT; ++q) z[q] = prox_elem<REG>(av[q], gl, blo[q], bhi[q]);     // :154
            if (lane == 0) {
                double *hp = hist_w + (k & (AD_HIST - 1)) * 4;
                hp[0] = gam; hp[1] = fi_z; hp[2] = cnew;
            }
            __syncwarp();
            // every CTA keeps its own copy of the (bit-identical) scalars, so that what its producer warp stages later was written
            // by this CTA and is ordered by the CTA-scope release/acquire chain through empty_bar
            if (tid == 0) {
                double2 *o = reinterpret_cast<double2 *>(p.ad + (size_t)rank * 4 * (size_t)p.N + 4 * (ik & CIAO_IDX_MASK));
                __stcg(o, make_double2(gam, fi_z));
                __stcg(o + 1, make_double2(cnew, 0.0));
            }
            ++done;
            PROF_T(t_f);
            PROF_ADD(5, t_e, t_f);  // main update, stores, history
            PROF_ADD(6, t_0, t_f);  // whole step
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            if (!valid[h]) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                p.v_z[gcol[h] + e] = z[2 * h + e];
                p.v_av[gcol[h] + e] = av[2 * h + e];
            }
        }
#ifdef CIAO_SEQ_PROFILE
        if (tid == 0)
            for (int i = 0; i < 8; ++i) g_seq_prof[rank * 8 + i] = prof_acc[i];
#endif
        if (rank == 0 && tid == 0) {
            p.scal[0] = hg;
            p.counters[0] = done;
            p.counters[1] = nbt;
        }
    }
    cluster_sync_all();
}

// ---------------------------------------------------------------------------
// Initialisation pass (Finito_adaptive.jl:65-87): one CTA per row at a time, two block reductions per row.
//   ad[i] = {γ_i, f_i(x0), c_i(x0), 0};  table row i = x0;  err = 1 if ∇f_i(x0 + 1) == ∇f_i(x0) (the reference's random
//   fallback :75-81 is not available on the device)
template <int LOSS>
__global__ void __launch_bounds__(256) adaptive_init_kernel(const double *rec, int64_t N, int64_t d, int64_t d_pad, int64_t ld,
                                                            const double *x0, double *table, double *ad, double alpha, double Nd,
                                                            int *err) {
    __shared__ double red[2][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t i = blockIdx.x; i < N; i += gridDim.x) {
        const double *row = rec + i * ld;
        double s0 = 0.0, s1 = 0.0;
        for (int64_t k = tid; k < d_pad; k += blockDim.x) {
            const double ak = row[k], xk = x0[k];
            s0 = fma(ak, xk, s0);
            s1 = fma(ak, __dadd_rn(xk, 1.0), s1);   // xeps = x0 .+ 1  (:73); padding columns hold a = 0
            table[i * d_pad + k] = xk;              // push!(s, copy(x0))  (:67)
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; }
        __syncthreads();
        double u0 = 0.0, u1 = 0.0;
        for (int w = 0; w < 8; ++w) { u0 += red[0][w]; u1 += red[1][w]; }
        __syncthreads();
        const double b = row[d_pad + TAIL_B], lam = row[d_pad + TAIL_LAM];
        const double c0 = loss_coef<LOSS>(u0, b, lam), c1 = loss_coef<LOSS>(u1, b, lam);
        double q = 0.0;
        for (int64_t k = tid; k < d_pad; k += blockDim.x) {
            const double ak = row[k];
            const double df = __dsub_rn(grad_elem<LOSS>(ak, c1, lam), grad_elem<LOSS>(ak, c0, lam));
            q = fma(df, df, q);
        }
        q = warp_sum(q);
        if (lane == 0) red[0][warp] = q;
        __syncthreads();
        double nm2 = 0.0;
        for (int w = 0; w < 8; ++w) nm2 += red[0][w];
        __syncthreads();
        if (tid == 0) {
            const double nmg = __dsqrt_rn(nm2);                                               // :75
            if (nmg < 2.220446049250313e-16) atomicExch(err, 2);                              // :77: the host runs the random restart
            double L_int = __ddiv_rn(nmg, __dsqrt_rn((double)d));                             // :84 (t = 1)
            L_int = __ddiv_rn(L_int, Nd);                                                     // :85
            double2 *o = reinterpret_cast<double2 *>(ad + 4 * i);
            o[0] = make_double2(__ddiv_rn(alpha, L_int), loss_value<LOSS>(u0, b, lam));       // :86, :66
            o[1] = make_double2(c0, nmg);   // slot 3: ‖∇f_i(x0+1) − ∇f_i(x0)‖ for the host (the step kernel resets it to 0)
        }
    }
}

// one degenerate component: out[0] = ‖∇f_i(xeps) − ∇f_i(x0)‖ for a host-drawn xeps (the random restart, :79-81).  One CTA.
template <int LOSS>
__global__ void __launch_bounds__(256) adaptive_retry_kernel(const double *row, int64_t d_pad, const double *x0, const double *xeps,
                                                             double *out) {
    __shared__ double red[2][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double s0 = 0.0, s1 = 0.0;
    for (int64_t k = tid; k < d_pad; k += blockDim.x) {
        s0 = fma(row[k], x0[k], s0);
        s1 = fma(row[k], xeps[k], s1);
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; }
    __syncthreads();
    double u0 = 0.0, u1 = 0.0;
    for (int w = 0; w < 8; ++w) { u0 += red[0][w]; u1 += red[1][w]; }
    __syncthreads();
    const double b = row[d_pad + TAIL_B], lam = row[d_pad + TAIL_LAM];
    const double c0 = loss_coef<LOSS>(u0, b, lam), c1 = loss_coef<LOSS>(u1, b, lam);
    double q = 0.0;
    for (int64_t k = tid; k < d_pad; k += blockDim.x) {
        const double df = __dsub_rn(grad_elem<LOSS>(row[k], c1, lam), grad_elem<LOSS>(row[k], c0, lam));
        q = fma(df, df, q);
    }
    q = warp_sum(q);
    if (lane == 0) red[0][warp] = q;
    __syncthreads();
    if (tid == 0) {
        double nm2 = 0.0;
        for (int w = 0; w < 8; ++w) nm2 += red[0][w];
        out[0] = __dsqrt_rn(nm2);
    }
}

// ws[blockIdx.x][d_pad] = Σ_{i in chunk} x0 ./ γ_i     (sum(s ./ γ) with s_i = x0, :90)
__global__ void __launch_bounds__(256) adaptive_sdivg_kernel(const double *x0, const double *ad, int64_t N, int64_t d_pad, double *ws) {
    const int64_t col = blockIdx.y * (int64_t)blockDim.x + threadIdx.x;
    if (col >= d_pad) return;
    const double xk = x0[col];
    double acc = 0.0;
    for (int64_t i = blockIdx.x; i < N; i += gridDim.x) acc += __ddiv_rn(xk, __ldg(ad + 4 * i));
    ws[(size_t)blockIdx.x * d_pad + col] = acc;
}

// av = γ̂·(S − G/N)   (:90)
__global__ void adaptive_av_kernel(const double *S, const double *G, double hat_gamma, double Nd, int64_t d_pad, double *av) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < d_pad) av[j] = __dmul_rn(hat_gamma, __dsub_rn(S[j], __ddiv_rn(G[j], Nd)));
}

// ---------------------------------------------------------------------------
template <int CPT, int LOSS, int REG>
static int launch_adaptive(ciao_ctx *c, const AdArgs &a, const SeqShape &sh) {
    auto kern = adaptive_kernel<CPT, LOSS, REG>;
    const int W = sh.Tc / 32;
    const size_t smem = (size_t)AD_D * (2 * (size_t)sh.Tc * CPT + AD_EXTRA) * 8 + 2 * (size_t)sh.npart_pad * 4 * 8 +
                        (size_t)W * AD_HIST * 4 * 8 + (2 * AD_D + 2) * 8 + 128;
    if (smem > 200 * 1024) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "adaptive Finito: d = %lld needs %zu bytes of shared memory", (long long)c->d, smem);
    static size_t configured[CIAO_MAX_DEVICES] = {};
    if (smem > configured[c->device % CIAO_MAX_DEVICES]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[c->device % CIAO_MAX_DEVICES] = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sh.C);
    cfg.blockDim = dim3(sh.Tc + 32);
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

template <int CPT, int LOSS>
static int launch_adaptive_reg(ciao_ctx *c, const AdArgs &a, const SeqShape &sh) {
    switch (c->reg.kind) {
        case CIAO_REG_NORML1: return launch_adaptive<CPT, LOSS, CIAO_REG_NORML1>(c, a, sh);
        case CIAO_REG_INDBOX: return launch_adaptive<CPT, LOSS, CIAO_REG_INDBOX>(c, a, sh);
        default: return launch_adaptive<CPT, LOSS, CIAO_REG_ZERO>(c, a, sh);
    }
}

template <int CPT>
static int launch_adaptive_loss(ciao_ctx *c, const AdArgs &a, const SeqShape &sh) {
    return c->loss_kind == CIAO_LOSS_LS ? launch_adaptive_reg<CPT, CIAO_LOSS_LS>(c, a, sh)
                                        : launch_adaptive_reg<CPT, CIAO_LOSS_LOGISTIC>(c, a, sh);
}

// K steps on prepared indices; c->adapt_counters[0..1] = {steps completed, reductions of γ} afterwards
int run_seq_adaptive(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, double alpha, double tol_b) {
    NvtxRange nvtx("ciao:seq:finito_adaptive");
    if (K <= 0) return CIAO_OK;
    if (K >= (int64_t)1 << 31) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "adaptive Finito: more than 2^31 - 1 steps in one call");
    SeqShape sh;
    CIAO_TRY(seq_shape(c, &sh));
    if (sh.C > 8) sh.C = 8, sh.dc = c->d_pad / 8;
    if (sh.cpt > 4) {  // four register arrays of CPT doubles per thread: keep CPT ≤ 4 by widening the CTA
        const int64_t T = ((sh.dc + 3) / 4 + 31) / 32 * 32;
        if (T > 256) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "adaptive Finito: d = %lld too large", (long long)c->d);
        sh.cpt = 4; sh.Tc = (int)T;
    }
    sh.npart_pad = (sh.C * (sh.Tc / 32) + 31) / 32 * 32;
    AdArgs a;
    a.rec = c->rec; a.ld = c->ld; a.d_pad = c->d_pad; a.dc = sh.dc; a.idx = idx_prepared; a.K = K;
    a.table = c->table; a.ad = c->adapt; a.N = c->N_total; a.scal = c->adapt_scal; a.counters = c->adapt_counters;
    a.v_z = ctx_vec(c, CIAO_VEC_Z); a.v_av = ctx_vec(c, CIAO_VEC_AV);
    a.Nd = (double)c->N_total; a.alpha = alpha; a.tol_b = tol_b; a.npart_pad = sh.npart_pad; a.zero = 0; a.reg = c->reg;
    CUDA_TRY(cudaEventRecord(c->ev_sa, c->stream));
    int rc;
    switch (sh.cpt) {
        case 2: rc = launch_adaptive_loss<2>(c, a, sh); break;
        default: rc = launch_adaptive_loss<4>(c, a, sh); break;
    }
    CIAO_TRY(rc);
    CUDA_TRY(cudaEventRecord(c->ev_sb, c->stream));
    c->timing.launches += 1;
    c->timing.last_seq_steps = K;
    c->seq_timed = true;
    return CIAO_OK;
}

// table and per-component scalars {γ_i (t = 1), f_i(x0), c_i(x0), ‖∇f_i(x0+1) − ∇f_i(x0)‖}; the error flag is 2 if some norm is < eps
int run_adaptive_init(ciao_ctx *c, const double *x0_dev, double alpha) {
    const int grid = (int)std::min<int64_t>(c->N_total, (int64_t)c->num_sms * 8);
    if (c->loss_kind == CIAO_LOSS_LS)
        adaptive_init_kernel<CIAO_LOSS_LS><<<grid, 256, 0, c->stream>>>(c->rec, c->N_total, c->d, c->d_pad, c->ld, x0_dev, c->table, c->adapt,
                                                                       alpha, (double)c->N_total, c->err_dev);
    else
        adaptive_init_kernel<CIAO_LOSS_LOGISTIC><<<grid, 256, 0, c->stream>>>(c->rec, c->N_total, c->d, c->d_pad, c->ld, x0_dev, c->table,
                                                                             c->adapt, alpha, (double)c->N_total, c->err_dev);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

// ‖∇f_i(xeps) − ∇f_i(x0)‖ for component i (0-based) → out_dev[0]
int run_adaptive_retry(ciao_ctx *c, int64_t i, const double *x0_dev, const double *xeps_dev, double *out_dev) {
    const double *row = c->rec + i * c->ld;
    if (c->loss_kind == CIAO_LOSS_LS) adaptive_retry_kernel<CIAO_LOSS_LS><<<1, 256, 0, c->stream>>>(row, c->d_pad, x0_dev, xeps_dev, out_dev);
    else adaptive_retry_kernel<CIAO_LOSS_LOGISTIC><<<1, 256, 0, c->stream>>>(row, c->d_pad, x0_dev, xeps_dev, out_dev);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

// Σ_i x0/γ_i as chunk partials in c->ws (reduced into c->partial by the caller's reduce)
int run_adaptive_sdivg(ciao_ctx *c, const double *x0_dev, int *n_chunks_out) {
    const int chunks_y = (int)((c->d_pad + 255) / 256);
    int rows = std::max(1, c->num_sms * 8 / chunks_y);
    if (rows > c->N_total) rows = (int)c->N_total;
    const size_t need = ((size_t)rows * c->d_pad + rows + 16) * sizeof(double);
    if (need > c->ws_bytes) {
        if (c->ws) cudaFree(c->ws);
        c->ws = nullptr;
        c->ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->ws, need));
        c->ws_bytes = need;
    }
    adaptive_sdivg_kernel<<<dim3(rows, chunks_y), 256, 0, c->stream>>>(x0_dev, c->adapt, c->N_total, c->d_pad, c->ws);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    *n_chunks_out = rows;
    return CIAO_OK;
}

int run_adaptive_av(ciao_ctx *c, const double *S_dev, const double *G_dev, double hat_gamma) {
    adaptive_av_kernel<<<(int)((c->d_pad + 255) / 256), 256, 0, c->stream>>>(S_dev, G_dev, hat_gamma, (double)c->N_total, c->d_pad,
                                                                            ctx_vec(c, CIAO_VEC_AV));
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

#ifdef CIAO_SEQ_PROFILE
extern "C" int ciao_debug_seq_prof_adaptive(ciao_ctx *c, long long *out128) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpyFromSymbol(out128, g_seq_prof, 16 * 8 * sizeof(long long)));
    return CIAO_OK;
}
#endif

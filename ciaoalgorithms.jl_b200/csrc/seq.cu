// seq.cu — K3..K6: the sequential inner loops as ONE persistent thread-block-cluster
// kernel per call (no kernel launch per sample).
//
//   ALG_SVRG    SVRG_basic.jl:73-87      m inner steps + snapshot average
//   ALG_SAGA    SAGA_basic.jl:55-65      K single-sample steps with the N×d gradient table (SAGA and SAG)
//   ALG_FINITO  Finito_basic.jl:110-118  steps/batches with the N×d table s_i and per-component γ_i
//   ALG_LFINITO Finito_LFinito.jl:91-100 one sweep of corrections over all batches (no table)
//
// Design.  A cluster of C CTAs splits the d columns; a thread
// owns CPT fixed columns of every state vector (w/z, av, z_full, Σw) in REGISTERS
// for the whole call.  Per step:
//   1. the sampled row slice a_i[cols of this CTA] + its scalars (b_i, λ_i, γ_i) were
//      TMA-prefetched (cp.async.bulk, D-deep mbarrier ring) from the host-generated
//      index sequence, D−1 steps ahead;
//   2. partial dots → warp shuffles → each warp pushes its partial into EVERY CTA of
//      the cluster through DSMEM with st.async (remote store that completes on the
//      destination CTA's mbarrier by tx-count, so neither side needs a cluster-scope
//      fence); every thread then sums the C·W partials in one fixed order, so all
//      CTAs hold bit-identical scalars — no CTA or cluster barrier instruction on
//      the step's critical path;
//   3. the variance-reduced / aggregated update, the table row write and prox_g are
//      fused, element by element, in the reference's rounding order.
// Table rows are prefetched P steps ahead into registers by their owner threads
// (generic proxy, same thread reads and writes a given address → coherent); when a
// row index repeats inside the prefetch window the step is flagged by
// prep_indices_kernel and reloads the row after the previous write.
#include <algorithm>

#include "common.cuh"

enum { ALG_SVRG = 1, ALG_SAGA = 2, ALG_FINITO = 3, ALG_LFINITO = 4 };

struct SeqArgs {
    const double *rec;
    int64_t ld, d_pad, dc;  // dc = columns per CTA
    const int64_t *idx;     // prepared: 0-based row | flags
    int64_t K;
    double *table;
    double *v_z, *v_zfull, *v_w, *v_av, *v_zsum;
    double gamma, hat_gamma, Nd, m_d;
    int plus, sag;
    RegParams reg;
};

#ifdef CIAO_SEQ_PROFILE
__device__ long long g_seq_prof[4];  // accumulated cycles of thread 0 of CTA 0 per phase (debug builds only)
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(i, a, b) prof_acc[i] += (b) - (a)
#else
#define PROF_T(var)
#define PROF_ADD(i, a, b)
#endif

#define SEQ_MAX_PART 128  // C·W ≤ 128

__device__ __forceinline__ double prox_rt(int kind, double x, double gl, double lo, double hi) {
    if (kind == CIAO_REG_NORML1) return prox_elem<CIAO_REG_NORML1>(x, gl, lo, hi);
    if (kind == CIAO_REG_INDBOX) return prox_elem<CIAO_REG_INDBOX>(x, gl, lo, hi);
    return x;
}

template <int CPT, int ALG, int LOSS>
__global__ void __launch_bounds__(512, 1) seq_kernel(const SeqArgs p) {
    constexpr bool TABLE = (ALG == ALG_SAGA || ALG == ALG_FINITO);
    constexpr bool TWO_DOTS = (ALG == ALG_SVRG || ALG == ALG_LFINITO);
    constexpr int P = CPT >= 8 ? 4 : 6;  // prefetch distance in steps
    constexpr int D = P + 1;             // row ring depth
    constexpr int H = CPT / 2;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const uint32_t rank = cluster_ctarank(), C = cluster_nctarank();
    const int64_t dc = p.dc;
    const size_t slot_doubles = (size_t)dc + CIAO_TAIL;
    double *ring = reinterpret_cast<double *>(smem_raw);
    double *part = ring + D * slot_doubles;                 // [2][SEQ_MAX_PART][2]
    uint64_t *row_bar = reinterpret_cast<uint64_t *>(part + 2 * SEQ_MAX_PART * 2);
    uint64_t *part_bar = row_bar + D;

    if (tid == 0) {
        for (int s = 0; s < D; ++s) mbar_init(&row_bar[s], 1);
        mbar_init(&part_bar[0], 1);  // one local arrive (expect_tx) per phase; the data arrives as tx bytes
        mbar_init(&part_bar[1], 1);
        fence_mbar_init();
    }
    cluster_sync_all();  // peers' barriers exist before anyone arrives on them

    const int64_t K = p.K;
    const int64_t cbase = (int64_t)rank * dc;
    auto issue_row = [&](int64_t step, int64_t pidx) {
        const int64_t i = pidx & CIAO_IDX_MASK;
        const int slot = (int)(step % D);
        double *dst = ring + slot * slot_doubles;
        const double *src = p.rec + i * p.ld;
        mbar_arrive_expect_tx(&row_bar[slot], (uint32_t)(dc * 8 + CIAO_TAIL * 8));
        tma_load_1d(dst, src + cbase, (uint32_t)(dc * 8), &row_bar[slot]);
        tma_load_1d(dst + dc, src + p.d_pad, CIAO_TAIL * 8, &row_bar[slot]);
    };

    // ---- this thread's columns and state registers -------------------------------
    int lcol[H];
    int64_t gcol[H];
    double z[CPT], av[CPT], zf[CPT], zs[CPT], blo[CPT], bhi[CPT];
    const int rk = p.reg.kind;
#pragma unroll
    for (int h = 0; h < H; ++h) {
        lcol[h] = 2 * (tid + T * h);
        if (lcol[h] >= dc) lcol[h] = -1;
        gcol[h] = cbase + (lcol[h] < 0 ? 0 : lcol[h]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int q = 2 * h + e;
            const bool v = lcol[h] >= 0;
            const int64_t g = gcol[h] + e;
            // the running iterate: SVRG → w, others → z
            z[q] = v ? (ALG == ALG_SVRG ? p.v_w[g] : p.v_z[g]) : 0.0;
            av[q] = v ? p.v_av[g] : 0.0;
            zf[q] = (v && TWO_DOTS) ? p.v_zfull[g] : 0.0;
            zs[q] = (v && ALG == ALG_SVRG) ? p.v_zsum[g] : 0.0;
            blo[q] = (v && p.reg.lo_v) ? p.reg.lo_v[g] : p.reg.lo_s;
            bhi[q] = (v && p.reg.hi_v) ? p.reg.hi_v[g] : p.reg.hi_s;
        }
    }
    const double gstep = (ALG == ALG_SVRG || ALG == ALG_SAGA) ? p.gamma : p.hat_gamma;
    const double gl = gstep * p.reg.lambda;
    const double cN = __ddiv_rn(p.hat_gamma, p.Nd);  // LFinito: γ̂/N

    // ---- index / table-row prefetch pipelines -------------------------------------
    int64_t iq[P];
    double2 tbuf[P][H];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        iq[j] = (j < K) ? __ldg(p.idx + j) : 0;
        if (TABLE && j < K) {
            const double *trow = p.table + (iq[j] & CIAO_IDX_MASK) * p.d_pad;
#pragma unroll
            for (int h = 0; h < H; ++h)
                tbuf[j][h] = (lcol[h] >= 0) ? __ldcg(reinterpret_cast<const double2 *>(trow + gcol[h])) : make_double2(0, 0);
        }
    }
    int64_t in1 = (P < K) ? __ldg(p.idx + P) : 0;          // index of step k+P   (k = 0)
    int64_t in2 = (P + 1 < K) ? __ldg(p.idx + P + 1) : 0;  // index of step k+P+1
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < P; ++j)
            if (j < K) issue_row(j, iq[j]);
        if (P < K) issue_row(P, in1);
    }

    const uint32_t part_bytes = C * W * 16;
#ifdef CIAO_SEQ_PROFILE
    long long prof_acc[4] = {0, 0, 0, 0};
#endif
    int slot = 0;
    uint32_t row_phase = 0;
    for (int64_t k0 = 0; k0 < K; k0 += P) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int64_t k = k0 + j;
            if (k >= K) break;
            const int par = (int)(k & 1);
            if (tid == 0) {
                mbar_arrive_expect_tx(&part_bar[par], part_bytes);  // arm this step's exchange phase
                // slot (k−1)%D was fully read by every warp before it sent its partial for step k−1
                if (k >= 1 && k + P < K) issue_row(k + P, in1);
            }
            const int64_t ik = iq[j];
            PROF_T(t_a);
            mbar_wait(&row_bar[slot], row_phase);
            PROF_T(t_b);
            const double *rp = ring + slot * slot_doubles;
            double a[CPT];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                double2 v = make_double2(0.0, 0.0);
                if (lcol[h] >= 0) v = *reinterpret_cast<const double2 *>(rp + lcol[h]);
                a[2 * h] = v.x;
                a[2 * h + 1] = v.y;
            }
            const double tb = rp[dc], tl = rp[dc + 1], tgam = rp[dc + 2];

            if (ALG == ALG_LFINITO && (ik & CIAO_FLAG_PROX)) {  // Finito_LFinito.jl:92
#pragma unroll
                for (int q = 0; q < CPT; ++q) z[q] = prox_rt(rk, av[q], gl, blo[q], bhi[q]);
            }

            // ---- dots: v0 = a·(w|z), v1 = a·z_full ------------------------------------
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int q = 0; q < CPT; ++q) {
                v0 = fma(a[q], z[q], v0);
                if (TWO_DOTS) v1 = fma(a[q], zf[q], v1);
            }
            v0 = warp_sum(v0);
            if (TWO_DOTS) v1 = warp_sum(v1);
            PROF_T(t_c);
            if (lane < C) {
                const uint32_t slot_addr = smem_u32(part + ((size_t)par * SEQ_MAX_PART + rank * W + warp) * 2);
                st_async_v2f64(mapa_u32(slot_addr, lane), v0, v1, mapa_u32(smem_u32(&part_bar[par]), lane));
            }
            mbar_wait(&part_bar[par], (uint32_t)((k >> 1) & 1));
            PROF_T(t_d);
            double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
            {
                const double2 *pp = reinterpret_cast<const double2 *>(part + (size_t)par * SEQ_MAX_PART * 2);
                const int n = C * W;
                for (int e = 0; e < n; e += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (e + u < n) {
                            const double2 v = pp[e + u];
                            s0[u] += v.x;
                            if (TWO_DOTS) s1[u] += v.y;
                        }
                }
            }
            const double u0 = (s0[0] + s0[1]) + (s0[2] + s0[3]);
            const double u1 = (s1[0] + s1[1]) + (s1[2] + s1[3]);

            // ---- fused update ---------------------------------------------------------
            if (ALG == ALG_SVRG) {  // SVRG_basic.jl:74-81
                const double cz = loss_coef<LOSS>(u1, tb, tl), cw = loss_coef<LOSS>(u0, tb, tl);
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
                    double t = __dsub_rn(grad_elem<LOSS>(a[q], cz, tl), grad_elem<LOSS>(a[q], cw, tl));
                    t = __dsub_rn(t, av[q]);
                    t = __dmul_rn(t, p.gamma);
                    t = __dadd_rn(t, z[q]);
                    z[q] = prox_rt(rk, t, gl, blo[q], bhi[q]);
                    zs[q] = __dadd_rn(zs[q], z[q]);
                }
            } else if (ALG == ALG_LFINITO) {  // Finito_LFinito.jl:94-98
                const double czf = loss_coef<LOSS>(u1, tb, tl), czz = loss_coef<LOSS>(u0, tb, tl);
                const double rr = __ddiv_rn(p.hat_gamma, tgam);
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
                    av[q] = __dadd_rn(av[q], __dmul_rn(cN, grad_elem<LOSS>(a[q], czf, tl)));
                    av[q] = __dsub_rn(av[q], __dmul_rn(cN, grad_elem<LOSS>(a[q], czz, tl)));
                    av[q] = __dadd_rn(av[q], __dmul_rn(rr, __dsub_rn(z[q], zf[q])));
                }
            } else {
                const double c = loss_coef<LOSS>(u0, tb, tl);
                double2 sold[H];
                double *trow = p.table + (ik & CIAO_IDX_MASK) * p.d_pad;
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    sold[h] = tbuf[j][h];
                    if ((ik & CIAO_FLAG_HAZARD) && lcol[h] >= 0)
                        sold[h] = __ldcg(reinterpret_cast<const double2 *>(trow + gcol[h]));
                }
                double snew[CPT];
                if (ALG == ALG_SAGA) {  // SAGA_basic.jl:56-65
#pragma unroll
                    for (int q = 0; q < CPT; ++q) {
                        const double so = (q & 1) ? sold[q / 2].y : sold[q / 2].x;
                        const double g = grad_elem<LOSS>(a[q], c, tl);
                        const double diff = __dsub_rn(g, so);
                        double w;
                        if (p.sag) {
                            av[q] = __dadd_rn(av[q], __ddiv_rn(diff, p.Nd));
                            w = __dsub_rn(z[q], __dmul_rn(p.gamma, av[q]));
                        } else {
                            w = __dsub_rn(z[q], __dmul_rn(p.gamma, __dadd_rn(diff, av[q])));
                            av[q] = __dadd_rn(av[q], __ddiv_rn(diff, p.Nd));
                        }
                        z[q] = prox_rt(rk, w, gl, blo[q], bhi[q]);
                        snew[q] = g;
                    }
                } else {  // Finito_basic.jl:112-118
                    const double cneg = -__ddiv_rn(tgam, p.Nd);
                    const double rr = __ddiv_rn(p.hat_gamma, tgam);
#pragma unroll
                    for (int q = 0; q < CPT; ++q) {
                        const double so = (q & 1) ? sold[q / 2].y : sold[q / 2].x;
                        double t = __dmul_rn(grad_elem<LOSS>(a[q], c, tl), cneg);
                        t = __dadd_rn(t, z[q]);
                        av[q] = __dadd_rn(av[q], __dmul_rn(__dsub_rn(t, so), rr));
                        snew[q] = t;
                    }
                    if (ik & CIAO_FLAG_PROX) {
#pragma unroll
                        for (int q = 0; q < CPT; ++q) z[q] = prox_rt(rk, av[q], gl, blo[q], bhi[q]);
                    }
                }
#pragma unroll
                for (int h = 0; h < H; ++h)
                    if (lcol[h] >= 0)
                        __stcg(reinterpret_cast<double2 *>(trow + gcol[h]), make_double2(snew[2 * h], snew[2 * h + 1]));
            }

            // ---- rotate the pipelines: this register slot now serves step k+P ---------
            iq[j] = in1;
            if (TABLE && k + P < K) {
                const double *nrow = p.table + (in1 & CIAO_IDX_MASK) * p.d_pad;
#pragma unroll
                for (int h = 0; h < H; ++h)
                    if (lcol[h] >= 0) tbuf[j][h] = __ldcg(reinterpret_cast<const double2 *>(nrow + gcol[h]));
            }
            in1 = in2;
            in2 = (k + P + 2 < K) ? __ldg(p.idx + k + P + 2) : 0;
            PROF_T(t_e);
            PROF_ADD(0, t_a, t_b);  // wait for the prefetched row
            PROF_ADD(1, t_b, t_c);  // LDS + dots + warp shuffles
            PROF_ADD(2, t_c, t_d);  // cluster exchange
            PROF_ADD(3, t_d, t_e);  // sum of partials + fused update + pipeline rotation
            if (++slot == D) {
                slot = 0;
                row_phase ^= 1;
            }
        }
    }

    // ---- epilogue: state back to HBM ------------------------------------------------
#pragma unroll
    for (int h = 0; h < H; ++h) {
        if (lcol[h] < 0) continue;
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
    if (tid == 0 && rank == 0)
        for (int i = 0; i < 4; ++i) g_seq_prof[i] = prof_acc[i];
#endif
    cluster_sync_all();  // nobody exits while a peer may still touch its shared memory
}

// ---------------------------------------------------------------------------
template <int CPT, int ALG, int LOSS>
static int launch_seq(ciao_ctx *c, const SeqArgs &a, int C, int T) {
    auto kern = seq_kernel<CPT, ALG, LOSS>;
    constexpr int P = CPT >= 8 ? 4 : 6;
    const size_t smem = (size_t)(P + 1) * (a.dc + CIAO_TAIL) * 8 + 2 * SEQ_MAX_PART * 2 * 8 + (P + 1 + 2) * 8 + 128;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    return CIAO_OK;
}

template <int CPT, int ALG>
static int launch_seq_loss(ciao_ctx *c, const SeqArgs &a, int C, int T) {
    return c->loss_kind == CIAO_LOSS_LS ? launch_seq<CPT, ALG, CIAO_LOSS_LS>(c, a, C, T)
                                        : launch_seq<CPT, ALG, CIAO_LOSS_LOGISTIC>(c, a, C, T);
}

template <int ALG>
static int launch_seq_alg(ciao_ctx *c, const SeqArgs &a, int C, int T, int cpt) {
    switch (cpt) {
        case 2: return launch_seq_loss<2, ALG>(c, a, C, T);
        case 4: return launch_seq_loss<4, ALG>(c, a, C, T);
        default: return launch_seq_loss<8, ALG>(c, a, C, T);
    }
}

// Chooses the cluster shape: C CTAs × T threads × CPT columns per thread cover d_pad.
static int seq_shape(ciao_ctx *c, int *C_out, int *T_out, int *cpt_out, int64_t *dc_out) {
    const int64_t d_pad = c->d_pad;
    int C = c->seq_cluster > 0 ? c->seq_cluster : (d_pad >= 2048 ? 8 : (d_pad >= 512 ? 4 : 1));
    while (C > 1 && (d_pad % (4 * C) != 0)) C >>= 1;
    const int64_t dc = d_pad / C;
    const int T_target = c->seq_threads > 0 ? c->seq_threads : 128;
    int cpt = 2;
    while (cpt < 8 && (dc + cpt - 1) / cpt > T_target) cpt *= 2;
    const int64_t T = ((dc + cpt - 1) / cpt + 31) / 32 * 32;
    if (T > 512 || C * (T / 32) > SEQ_MAX_PART)
        CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "sequential kernel: d = %lld too large for cluster %d", (long long)c->d, C);
    *C_out = C; *T_out = (int)T; *cpt_out = cpt; *dc_out = dc;
    return CIAO_OK;
}

int run_seq(ciao_ctx *c, int alg, const int64_t *idx_prepared, int64_t K, double m_d) {
    if (K <= 0) return CIAO_OK;
    int C, T, cpt;
    int64_t dc;
    CIAO_TRY(seq_shape(c, &C, &T, &cpt, &dc));
    SeqArgs a;
    a.rec = c->rec; a.ld = c->ld; a.d_pad = c->d_pad; a.dc = dc;
    a.idx = idx_prepared; a.K = K; a.table = c->table;
    a.v_z = ctx_vec(c, CIAO_VEC_Z); a.v_zfull = ctx_vec(c, CIAO_VEC_Z_FULL); a.v_w = ctx_vec(c, CIAO_VEC_W);
    a.v_av = ctx_vec(c, CIAO_VEC_AV); a.v_zsum = ctx_vec(c, CIAO_VEC_Z);  // SVRG: state.z is the running sum of inner iterates
    a.gamma = c->gamma; a.hat_gamma = c->hat_gamma; a.Nd = (double)c->N_total; a.m_d = m_d;
    a.plus = c->plus; a.sag = c->sag; a.reg = c->reg;
    CUDA_TRY(cudaEventRecord(c->ev_sa, c->stream));
    int rc;
    switch (alg) {
        case ALG_SVRG: rc = launch_seq_alg<ALG_SVRG>(c, a, C, T, cpt); break;
        case ALG_SAGA: rc = launch_seq_alg<ALG_SAGA>(c, a, C, T, cpt); break;
        case ALG_FINITO: rc = launch_seq_alg<ALG_FINITO>(c, a, C, T, cpt); break;
        default: rc = launch_seq_alg<ALG_LFINITO>(c, a, C, T, cpt); break;
    }
    CIAO_TRY(rc);
    CUDA_TRY(cudaEventRecord(c->ev_sb, c->stream));
    c->timing.launches += 1;
    c->timing.last_seq_steps = K;
    c->seq_timed = true;
    return CIAO_OK;
}

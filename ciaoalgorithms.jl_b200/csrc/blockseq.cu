// blockseq.cu — the sequential loops and table-init passes for components that are M×d BLOCKS
//   f_i = LeastSquares(A_i (M×d), b_i (M), λ_i)   or   Precompose(LogisticLoss(y_i (M), μ_i), L_i (M×d))
// ("general m_i×d blocks per f_i"; test_lasso.jl:52-54 builds the M = 1 case), and for the complex
// soft-threshold (CIAO_REG_NORML1_PAIRS): a complex 1×d LeastSquares row is the M = 2 real block [Re; Im] of the realified
// problem x = (re_0, im_0, re_1, im_1, …), so genuinely complex data (test_lasso.jl:3 with non-zero imaginary parts) runs here.
//
//   SVRG_basic.jl:73-87, SAGA_basic.jl:41-65, Finito_basic.jl:76-118, Finito_LFinito.jl:91-100 with ∇f_i = Σ_r c_r·a_r
//   (ProximalOperators 0.14, dense M×d: res = A x − b; y = Aᴴ res accumulated over the rows r = 1…M; y .*= λ).
//
// This is the GENERAL path, not the tuned one: one persistent CTA of 1024 threads, a thread owns the column pairs
// u = t, t + 1024, … of every state vector in registers, M row dots per step go through a two-stage block reduction
// (warp shuffles → shared memory → one warp per value), the rows are re-read from L2 for the gradient.  ≈ 2–4 µs per step
// against 0.3 µs of the cluster kernels (seq_impl.cuh), which handle M = 1 with the real soft-threshold; the arithmetic per
// element is the same, in the reference's rounding order.  No CPU fallback exists for these cases either.
#include <algorithm>

#include "common.cuh"

constexpr int BLK_T = 1024;      // threads of the CTA
constexpr int BLK_U = 4;         // column pairs per thread: d_pad ≤ 2·BLK_T·BLK_U = 8192
constexpr int BLK_RB = 4;        // rows per reduction round

struct BlkArgs {
    const double *rec;           // [N·M][ld]
    int64_t ld, d_pad;
    int M;
    const int64_t *idx;          // prepared: 0-based component | PROX flag
    int64_t K;
    double *table;               // [N][d_pad]
    const double *gam;           // [N] γ_i (Finito, LFinito)
    double *v_z, *v_zfull, *v_w, *v_av, *v_zsum;
    double gamma, hat_gamma, Nd, m_d;
    int plus, sag;
    RegParams reg;
    const int *err;
    // table init
    int64_t N;
    const double *x0;
    double *ws;                  // [grid][d_pad] partial Σ
    double *fws;
};

// prox_g on one (even, odd) column pair; CIAO_REG_NORML1_PAIRS: sign(x)·max(0, |x| − γλ) on the complex number (re, im)
__device__ __forceinline__ double2 prox_pair(const RegParams &reg, double2 x, double gl, int64_t col) {
    if (reg.kind == CIAO_REG_NORML1_PAIRS) {
        const double ab = hypot(x.x, x.y);
        const double m = ab - gl > 0 ? ab - gl : 0.0;
        return ab == 0 ? make_double2(0.0, 0.0) : make_double2(__dmul_rn(__ddiv_rn(x.x, ab), m), __dmul_rn(__ddiv_rn(x.y, ab), m));
    }
    const double lo0 = reg.lo_v ? reg.lo_v[col] : reg.lo_s, hi0 = reg.hi_v ? reg.hi_v[col] : reg.hi_s;
    const double lo1 = reg.lo_v ? reg.lo_v[col + 1] : reg.lo_s, hi1 = reg.hi_v ? reg.hi_v[col + 1] : reg.hi_s;
    return make_double2(prox_rt(reg.kind, x.x, gl, lo0, hi0), prox_rt(reg.kind, x.y, gl, lo1, hi1));
}

// c_r = ∇ℓ(a_r·x, b_r) for the M rows of component i, x held in registers (xs: 1 or 2 vectors) → coef[v][r] in shared memory.
// Two-stage block reduction, BLK_RB rows per round: warp shuffles, then warp q sums value q over the 32 warp partials.
template <int LOSS, int NV>
__device__ __forceinline__ void block_coefs(const BlkArgs &p, int64_t i, const double2 (&xa)[BLK_U], const double2 (&xb)[BLK_U],
                                            const bool (&valid)[BLK_U], double *red, double *coef) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = p.M;
    for (int r0 = 0; r0 < M; r0 += BLK_RB) {
        double pa[BLK_RB], pb[BLK_RB];
#pragma unroll
        for (int q = 0; q < BLK_RB; ++q) {
            pa[q] = pb[q] = 0.0;
            if (r0 + q < M) {
                const double *row = p.rec + (i * M + r0 + q) * p.ld;
#pragma unroll
                for (int u = 0; u < BLK_U; ++u)
                    if (valid[u]) {
                        const double2 a = *reinterpret_cast<const double2 *>(row + 2 * (tid + BLK_T * u));
                        pa[q] = fma(a.x, xa[u].x, pa[q]);
                        pa[q] = fma(a.y, xa[u].y, pa[q]);
                        if (NV == 2) {
                            pb[q] = fma(a.x, xb[u].x, pb[q]);
                            pb[q] = fma(a.y, xb[u].y, pb[q]);
                        }
                    }
            }
        }
#pragma unroll
        for (int q = 0; q < BLK_RB; ++q) {
            pa[q] = warp_sum(pa[q]);
            if (NV == 2) pb[q] = warp_sum(pb[q]);
            if (lane == 0) {
                red[(2 * q) * 32 + warp] = pa[q];
                red[(2 * q + 1) * 32 + warp] = pb[q];
            }
        }
        __syncthreads();
        if (warp < 2 * BLK_RB) {          // warp w sums value w (fixed order inside warp_sum → the same bits every time)
            const double s = warp_sum(red[warp * 32 + lane]);
            const int q = warp >> 1, v = warp & 1;
            if (lane == 0 && r0 + q < M && v < NV) {
                const double *row = p.rec + (i * M + r0 + q) * p.ld;
                coef[v * M + r0 + q] = loss_coef<LOSS>(s, row[p.d_pad + TAIL_B], row[p.d_pad + TAIL_LAM]);
            }
        }
        __syncthreads();
    }
}

// ∇f_i element pair from the coefficients: y = Σ_r a_r·c_r (rows in order, product and sum rounded separately), LS: y .*= λ
template <int LOSS>
__device__ __forceinline__ double2 block_grad_pair(const BlkArgs &p, int64_t i, int64_t col, const double *coef, double lam) {
    double y0 = 0.0, y1 = 0.0;
    for (int r = 0; r < p.M; ++r) {
        const double2 a = *reinterpret_cast<const double2 *>(p.rec + (i * p.M + r) * p.ld + col);
        const double c = coef[r];
        y0 = __dadd_rn(y0, __dmul_rn(a.x, c));
        y1 = __dadd_rn(y1, __dmul_rn(a.y, c));
    }
    if (LOSS == CIAO_LOSS_LS) {
        y0 = __dmul_rn(y0, lam);
        y1 = __dmul_rn(y1, lam);
    }
    return make_double2(y0, y1);
}

template <int ALG, int LOSS>
__global__ void __launch_bounds__(BLK_T, 1) block_seq_kernel(const BlkArgs p) {
    extern __shared__ double blk_sm[];
    if (*reinterpret_cast<const volatile int *>(p.err) != 0) return;   // an out-of-range index: no step runs
    double *red = blk_sm;                       // [2·BLK_RB][32]
    double *coef = red + 2 * BLK_RB * 32;       // [2][M]
    const int tid = threadIdx.x;
    constexpr bool USES_ZFULL = (ALG == ALG_SVRG || ALG == ALG_LFINITO);
    constexpr bool TABLE = (ALG == ALG_SAGA || ALG == ALG_FINITO);
    bool valid[BLK_U];
    int64_t col[BLK_U];
    double2 z[BLK_U], av[BLK_U], zf[BLK_U], zs[BLK_U];
#pragma unroll
    for (int u = 0; u < BLK_U; ++u) {
        col[u] = 2 * (int64_t)(tid + BLK_T * u);
        valid[u] = col[u] < p.d_pad;
        const int64_t c = valid[u] ? col[u] : 0;
        const double *zsrc = ALG == ALG_SVRG ? p.v_w : p.v_z;
        z[u] = valid[u] ? *reinterpret_cast<const double2 *>(zsrc + c) : make_double2(0.0, 0.0);
        av[u] = valid[u] ? *reinterpret_cast<const double2 *>(p.v_av + c) : make_double2(0.0, 0.0);
        zf[u] = (valid[u] && USES_ZFULL) ? *reinterpret_cast<const double2 *>(p.v_zfull + c) : make_double2(0.0, 0.0);
        zs[u] = (valid[u] && ALG == ALG_SVRG) ? *reinterpret_cast<const double2 *>(p.v_zsum + c) : make_double2(0.0, 0.0);
    }
    const double gstep = (ALG == ALG_SVRG || ALG == ALG_SAGA) ? p.gamma : p.hat_gamma;
    const double gl = gstep * p.reg.lambda;
    for (int64_t k = 0; k < p.K; ++k) {
        const int64_t ik = p.idx[k];
        const int64_t i = ik & CIAO_IDX_MASK;
        if (ALG == ALG_LFINITO && (ik & CIAO_FLAG_PROX)) {                               // Finito_LFinito.jl:92
#pragma unroll
            for (int u = 0; u < BLK_U; ++u)
                if (valid[u]) z[u] = prox_pair(p.reg, av[u], gl, col[u]);
        }
        block_coefs<LOSS, USES_ZFULL ? 2 : 1>(p, i, z, zf, valid, red, coef);
        const double lam = p.rec[(i * p.M) * p.ld + p.d_pad + TAIL_LAM];
        double gi = 0.0, gn = 0.0, hg = 0.0;
        if (ALG == ALG_FINITO || ALG == ALG_LFINITO) {
            gi = p.gam[i];
            gn = __ddiv_rn(gi, p.Nd);                // γ_i/N   Finito_basic.jl:113
            hg = __ddiv_rn(p.hat_gamma, gi);         // γ̂/γ_i   Finito_basic.jl:115
        }
        const double cN = __ddiv_rn(p.hat_gamma, p.Nd);
#pragma unroll
        for (int u = 0; u < BLK_U; ++u) {
            if (!valid[u]) continue;
            const double2 g = block_grad_pair<LOSS>(p, i, col[u], coef, lam);             // ∇f_i(w) / ∇f_i(z)
            if (ALG == ALG_SVRG) {                                                         // SVRG_basic.jl:74-81
                const double2 gz = block_grad_pair<LOSS>(p, i, col[u], coef + p.M, lam);   // ∇f_i(z_full)
                double2 t = make_double2(__dsub_rn(gz.x, g.x), __dsub_rn(gz.y, g.y));
                t = make_double2(__dsub_rn(t.x, av[u].x), __dsub_rn(t.y, av[u].y));
                t = make_double2(__dmul_rn(t.x, p.gamma), __dmul_rn(t.y, p.gamma));
                t = make_double2(__dadd_rn(t.x, z[u].x), __dadd_rn(t.y, z[u].y));
                z[u] = prox_pair(p.reg, t, gl, col[u]);
                zs[u] = make_double2(__dadd_rn(zs[u].x, z[u].x), __dadd_rn(zs[u].y, z[u].y));
            } else if (ALG == ALG_LFINITO) {                                               // Finito_LFinito.jl:94-98
                const double2 gz = block_grad_pair<LOSS>(p, i, col[u], coef + p.M, lam);
                av[u] = make_double2(__dadd_rn(av[u].x, __dmul_rn(cN, gz.x)), __dadd_rn(av[u].y, __dmul_rn(cN, gz.y)));
                av[u] = make_double2(__dsub_rn(av[u].x, __dmul_rn(cN, g.x)), __dsub_rn(av[u].y, __dmul_rn(cN, g.y)));
                av[u] = make_double2(__dadd_rn(av[u].x, __dmul_rn(hg, __dsub_rn(z[u].x, zf[u].x))),
                                     __dadd_rn(av[u].y, __dmul_rn(hg, __dsub_rn(z[u].y, zf[u].y))));
            } else {
                double2 *trow = reinterpret_cast<double2 *>(p.table + i * p.d_pad + col[u]);
                const double2 so = __ldcg(trow);
                double2 snew;
                if (ALG == ALG_SAGA) {                                                     // SAGA_basic.jl:56-65
                    const double2 diff = make_double2(__dsub_rn(g.x, so.x), __dsub_rn(g.y, so.y));
                    double2 w;
                    if (p.sag) {
                        av[u] = make_double2(__dadd_rn(av[u].x, __ddiv_rn(diff.x, p.Nd)), __dadd_rn(av[u].y, __ddiv_rn(diff.y, p.Nd)));
                        w = make_double2(__dsub_rn(z[u].x, __dmul_rn(p.gamma, av[u].x)), __dsub_rn(z[u].y, __dmul_rn(p.gamma, av[u].y)));
                    } else {
                        w = make_double2(__dsub_rn(z[u].x, __dmul_rn(p.gamma, __dadd_rn(diff.x, av[u].x))),
                                         __dsub_rn(z[u].y, __dmul_rn(p.gamma, __dadd_rn(diff.y, av[u].y))));
                        av[u] = make_double2(__dadd_rn(av[u].x, __ddiv_rn(diff.x, p.Nd)), __dadd_rn(av[u].y, __ddiv_rn(diff.y, p.Nd)));
                    }
                    z[u] = prox_pair(p.reg, w, gl, col[u]);
                    snew = g;
                } else {                                                                   // Finito_basic.jl:112-116
                    const double cneg = -gn;
                    const double2 t = make_double2(__dadd_rn(__dmul_rn(g.x, cneg), z[u].x), __dadd_rn(__dmul_rn(g.y, cneg), z[u].y));
                    av[u] = make_double2(__dadd_rn(av[u].x, __dmul_rn(__dsub_rn(t.x, so.x), hg)),
                                         __dadd_rn(av[u].y, __dmul_rn(__dsub_rn(t.y, so.y), hg)));
                    snew = t;
                }
                __stcg(trow, snew);      // read and written by this thread only: later steps see it in program order
            }
        }
        if (ALG == ALG_FINITO && (ik & CIAO_FLAG_PROX)) {                                  // Finito_basic.jl:118
#pragma unroll
            for (int u = 0; u < BLK_U; ++u)
                if (valid[u]) z[u] = prox_pair(p.reg, av[u], gl, col[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < BLK_U; ++u) {
        if (!valid[u]) continue;
        if (ALG == ALG_SVRG) {                                                             // SVRG_basic.jl:84-86
            const double2 zfull = make_double2(__ddiv_rn(zs[u].x, p.m_d), __ddiv_rn(zs[u].y, p.m_d));
            *reinterpret_cast<double2 *>(p.v_zfull + col[u]) = zfull;
            *reinterpret_cast<double2 *>(p.v_w + col[u]) = p.plus ? z[u] : zfull;
            *reinterpret_cast<double2 *>(p.v_zsum + col[u]) = make_double2(0.0, 0.0);
        } else {
            *reinterpret_cast<double2 *>(p.v_z + col[u]) = z[u];
            *reinterpret_cast<double2 *>(p.v_av + col[u]) = av[u];
        }
    }
}

// Table init for block components: MODE 1 (SAGA) s_i = ∇f_i(x0), Σ s_i;  MODE 2 (Finito) s_i = x0 − (γ_i/N)∇f_i(x0), Σ s_i/γ_i.
// CTA b takes components b, b + grid, …; its partial sums go to ws[b] and are closed by pass_tail_kernel.
template <int MODE, int LOSS>
__global__ void __launch_bounds__(BLK_T, 1) block_table_init_kernel(const BlkArgs p) {
    extern __shared__ double blk_sm[];
    double *red = blk_sm, *coef = red + 2 * BLK_RB * 32;
    const int tid = threadIdx.x;
    bool valid[BLK_U];
    int64_t col[BLK_U];
    double2 x[BLK_U], acc[BLK_U];
#pragma unroll
    for (int u = 0; u < BLK_U; ++u) {
        col[u] = 2 * (int64_t)(tid + BLK_T * u);
        valid[u] = col[u] < p.d_pad;
        x[u] = valid[u] ? *reinterpret_cast<const double2 *>(p.x0 + col[u]) : make_double2(0.0, 0.0);
        acc[u] = make_double2(0.0, 0.0);
    }
    double fsum = 0.0;
    for (int64_t i = blockIdx.x; i < p.N; i += gridDim.x) {
        block_coefs<LOSS, 1>(p, i, x, x, valid, red, coef);
        const double lam = p.rec[(i * p.M) * p.ld + p.d_pad + TAIL_LAM];
        if (tid == 0) {
            for (int r = 0; r < p.M; ++r) {   // Σ f_i for the objective: LS coef = residual, logistic value needs u — recompute from c is not possible, so LS only
                if (LOSS == CIAO_LOSS_LS) fsum += (lam / 2) * coef[r] * coef[r];
            }
        }
        const double gi = MODE == 2 ? p.gam[i] : 1.0;
        const double cg = MODE == 2 ? __ddiv_rn(gi, p.Nd) : 0.0;
#pragma unroll
        for (int u = 0; u < BLK_U; ++u) {
            if (!valid[u]) continue;
            double2 s = block_grad_pair<LOSS>(p, i, col[u], coef, lam);
            if (MODE == 2) {
                s = make_double2(__dsub_rn(x[u].x, __dmul_rn(cg, s.x)), __dsub_rn(x[u].y, __dmul_rn(cg, s.y)));   // Finito_basic.jl:79
                acc[u].x += __ddiv_rn(s.x, gi);
                acc[u].y += __ddiv_rn(s.y, gi);
            } else {
                acc[u].x += s.x;
                acc[u].y += s.y;
            }
            __stcs(reinterpret_cast<double2 *>(p.table + i * p.d_pad + col[u]), s);
        }
        __syncthreads();   // coef is rewritten by the next component
    }
#pragma unroll
    for (int u = 0; u < BLK_U; ++u)
        if (valid[u]) *reinterpret_cast<double2 *>(p.ws + (size_t)blockIdx.x * p.d_pad + col[u]) = acc[u];
    if (tid == 0) p.fws[blockIdx.x] = fsum;
}

// ---------------------------------------------------------------------------
static size_t blk_smem(const ciao_ctx *c) { return (size_t)(2 * BLK_RB * 32 + 2 * c->M + 8) * sizeof(double); }

static void blk_fill(ciao_ctx *c, BlkArgs &a) {
    a.rec = c->rec; a.ld = c->ld; a.d_pad = c->d_pad; a.M = c->M;
    a.table = c->table; a.gam = c->gamma_dev;
    a.v_z = ctx_vec(c, CIAO_VEC_Z); a.v_zfull = ctx_vec(c, CIAO_VEC_Z_FULL); a.v_w = ctx_vec(c, CIAO_VEC_W);
    a.v_av = ctx_vec(c, CIAO_VEC_AV); a.v_zsum = ctx_vec(c, CIAO_VEC_Z);
    a.gamma = c->gamma; a.hat_gamma = c->hat_gamma; a.Nd = (double)c->N_total;
    a.plus = c->plus; a.sag = c->sag; a.reg = c->reg; a.err = c->err_dev;
    a.N = c->N_total; a.idx = nullptr; a.K = 0; a.m_d = 1.0; a.x0 = nullptr; a.ws = nullptr; a.fws = nullptr;
}

template <int ALG>
static int launch_block_seq(ciao_ctx *c, const BlkArgs &a) {
    if (c->loss_kind == CIAO_LOSS_LS) block_seq_kernel<ALG, CIAO_LOSS_LS><<<1, BLK_T, blk_smem(c), c->stream>>>(a);
    else block_seq_kernel<ALG, CIAO_LOSS_LOGISTIC><<<1, BLK_T, blk_smem(c), c->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return CIAO_OK;
}

// K steps of `alg` on prepared indices through the general block kernel
int run_block_seq(ciao_ctx *c, int alg, const int64_t *idx_prepared, int64_t K, double m_d) {
    if (K <= 0) return CIAO_OK;
    if (c->d_pad > 2 * BLK_T * BLK_U) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "block components: d = %lld exceeds 8192", (long long)c->d);
    NvtxRange nvtx("ciao:seq:block_components");
    BlkArgs a;
    blk_fill(c, a);
    a.idx = idx_prepared; a.K = K; a.m_d = m_d;
    CUDA_TRY(cudaEventRecord(c->ev_sa, c->stream));
    int rc;
    switch (alg) {
        case ALG_SVRG: rc = launch_block_seq<ALG_SVRG>(c, a); break;
        case ALG_SAGA: rc = launch_block_seq<ALG_SAGA>(c, a); break;
        case ALG_FINITO: rc = launch_block_seq<ALG_FINITO>(c, a); break;
        default: rc = launch_block_seq<ALG_LFINITO>(c, a); break;
    }
    CIAO_TRY(rc);
    CUDA_TRY(cudaEventRecord(c->ev_sb, c->stream));
    c->timing.launches += 1;
    c->timing.last_seq_steps = K;
    c->seq_timed = true;
    return CIAO_OK;
}

// table init pass for block components: leaves the CTA partials in c->ws for the tail kernel; returns the grid
int run_block_table_init(ciao_ctx *c, int mode, const double *x0_dev, int *grid_out) {
    if (c->d_pad > 2 * BLK_T * BLK_U) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "block components: d = %lld exceeds 8192", (long long)c->d);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(c->N_total, c->num_sms));
    const size_t need = ((size_t)grid * c->d_pad + grid + 16) * sizeof(double);
    if (need > c->ws_bytes) {
        if (c->ws) cudaFree(c->ws);
        c->ws = nullptr;
        c->ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->ws, need));
        c->ws_bytes = need;
    }
    BlkArgs a;
    blk_fill(c, a);
    a.x0 = x0_dev; a.ws = c->ws; a.fws = c->ws + (size_t)grid * c->d_pad;
    const bool ls = c->loss_kind == CIAO_LOSS_LS;
    if (mode == PASS_SAGA_INIT) {
        if (ls) block_table_init_kernel<1, CIAO_LOSS_LS><<<grid, BLK_T, blk_smem(c), c->stream>>>(a);
        else block_table_init_kernel<1, CIAO_LOSS_LOGISTIC><<<grid, BLK_T, blk_smem(c), c->stream>>>(a);
    } else {
        if (ls) block_table_init_kernel<2, CIAO_LOSS_LS><<<grid, BLK_T, blk_smem(c), c->stream>>>(a);
        else block_table_init_kernel<2, CIAO_LOSS_LOGISTIC><<<grid, BLK_T, blk_smem(c), c->stream>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    *grid_out = grid;
    return CIAO_OK;
}

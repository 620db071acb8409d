/* ciao_oracle.c — CPU restatement of the CIAOAlgorithms.jl iteration loops.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the CPU
 * baseline for the B200 engine (libciao_cuda).  Nothing in the product path
 * (ciaoalgorithms.jl_b200/, include/) may include, link or call it; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs do.
 *
 * It restates, operation by operation and in the reference's order of
 * separate vector passes, the Julia loops of
 *   src/algorithms/SVRG/SVRG_basic.jl:30-96
 *   src/algorithms/SAGA_SAG/SAGA_basic.jl:26-68
 *   src/algorithms/Finito/Finito_basic.jl:44-121
 *   src/algorithms/Finito/Finito_LFinito.jl:40-103
 *   src/algorithms/Finito/Finito_adaptive.jl:59-160
 *   src/algorithms/ProShI/ProShI_basic.jl:44-132
 * together with the slice of the un-vendored dependency ProximalOperators.jl
 * 0.14 (Project.toml:10,16) those loops call: gradient!/prox! of
 * LeastSquares (dense 1×d), Precompose(LogisticLoss), Sum(Quadratic(diag),
 * SqrDistL2(IndBox)), NormL1, IndBox, Zero.
 *
 * Parity pinning: Julia is not installed in this image, so the reference
 * cannot be executed; the oracle is pinned against every golden vector the
 * reference's tests hold for this path (tests/test_oracle_golden.py):
 * x_star of test/test_logistic_l1.jl:29, sum_star of test/test_sharing.jl:28
 * and the planted optimum f* of test/test_lasso.jl:18-47, under the reference's
 * own maxit / stepsize choices and pass criteria (1e-4).  The reference holds
 * no per-step trajectories, so per-step parity rests on this restatement —
 * cross-checked after every step against a second, independent transcription
 * of the same Julia loops (tests/second_restatement.py, tests/test_oracle_cross.py:
 * agreement <= 4e-15 on every state vector and table row, all solver variants).
 *
 * Index sequences are INPUTS (1-based int64, exactly as Julia's rand / sample /
 * randperm would hand them over), never drawn here.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=x86-64-v3 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ciao_gen.h"

#define ORC_LOSS_LS 0       /* LeastSquares(A_i (1×d), b_i, λ_i)            */
#define ORC_LOSS_LOGISTIC 1 /* Precompose(LogisticLoss([y_i], μ_i), a_i', 1) */
#define ORC_LOSS_DIAGQUAD 2 /* Sum(Quadratic(diag(q_i), c_i), SqrDistL2(IndBox(lo,hi), η)) */

#define ORC_REG_ZERO 0
#define ORC_REG_NORML1 1
#define ORC_REG_INDBOX 2
#define ORC_REG_NORML1_PAIRS 3 /* NormL1(λ) on ComplexF64 data stored as interleaved (re, im) pairs: sign(x)·max(0, |x| − γλ) */

typedef struct {
    int32_t loss_kind;
    int32_t reg_kind;
    int64_t N;          /* number of components f_i                         */
    int64_t d;          /* dimension of x (rows) or of one block (sharing)  */
    int64_t lda;        /* leading dimension of A / Qdiag / qlin            */
    const double *A;    /* rows a_i (LS, logistic) or diag(Q_i) (diagquad)  */
    const double *b;    /* b_i (LS), y_i (logistic), linear terms N×d (diagquad) */
    const double *lam;  /* λ_i (LS) or μ_i (logistic), length N             */
    double box_lo, box_hi, eta; /* SqrDistL2(IndBox(lo,hi), η)              */
    double reg_lambda;  /* NormL1(λ)                                        */
    const double *reg_lo; /* IndBox lower bound: length d, or NULL → scalar */
    const double *reg_hi;
    double reg_lo_s, reg_hi_s;
    int64_t M;          /* rows per component: f_i = LeastSquares(A_i (M×d), b_i (M), λ_i) or Precompose(LogisticLoss(y_i (M)), L_i (M×d));
                           component i holds rows i·M … i·M+M−1 of A and b.  0 or 1: one row per component (test_lasso.jl:53).      */
} orc_problem;

/* ------------------------------------------------------------------------ */
/* ProximalOperators.jl 0.14 semantics                                       */
/* ------------------------------------------------------------------------ */

/* dot of a 1×d matrix row with x: BLAS gemv on a 1×d matrix; summation order
 * is BLAS-internal, restated here left to right. */
static double orc_dot(const double *a, const double *x, int64_t d) {
    double s = 0.0;
    for (int64_t k = 0; k < d; ++k) s += a[k] * x[k];
    return s;
}

/* gradient!(y, F[i], x); returns f_i(x).
 * LeastSquares (direct):  res = A x − b;  y = Aᴴ res;  y .*= λ;  f = (λ/2)·res²
 * Precompose(LogisticLoss): u = L x;  c = −μ y /(1 + exp(y u));  y = Lᴴ c;
 *                           f = μ·log(1 + 1/exp(y u))
 * Sum(Quadratic, SqrDistL2): y = Q x + q  +  η (x − Π_box x)
 */
double orc_gradient(const orc_problem *p, int64_t i, const double *x, double *y) {
    const int64_t d = p->d;
    if (p->M > 1 && p->loss_kind != ORC_LOSS_DIAGQUAD) {
        /* M×d blocks (ProximalOperators 0.14 leastSquaresDirect.jl / precompose.jl with a dense M×d matrix):
         *   LeastSquares:  res = A x − b (gemv);  y = Aᴴ res (gemv, accumulated over the rows r = 1…M);  y .*= λ;  f = (λ/2)‖res‖²
         *   Precompose(LogisticLoss(y, μ), L): u = L x;  c_r = −μ y_r/(1 + exp(y_r u_r));  y = Lᴴ c;  f = μ Σ_r log(1 + 1/exp(y_r u_r)) */
        const int64_t M = p->M;
        const double lam = p->lam[i];
        double val = 0.0;
        for (int64_t k = 0; k < d; ++k) y[k] = 0.0;
        for (int64_t r = 0; r < M; ++r) {
            const double *a = p->A + (i * M + r) * p->lda;
            const double br = p->b[i * M + r];
            double c;
            if (p->loss_kind == ORC_LOSS_LS) {
                c = orc_dot(a, x, d) - br;
                val += c * c;
            } else {
                double e = exp(br * orc_dot(a, x, d));
                c = -lam * br / (1 + e);
                val += log(1 + 1 / e);
            }
            for (int64_t k = 0; k < d; ++k) y[k] += a[k] * c;
        }
        if (p->loss_kind == ORC_LOSS_LS) {
            for (int64_t k = 0; k < d; ++k) y[k] *= lam;
            return (lam / 2) * val;
        }
        return lam * val;
    }
    const double *a = p->A + i * p->lda;
    if (p->loss_kind == ORC_LOSS_LS) {
        double res = orc_dot(a, x, d) - p->b[i];
        double lam = p->lam[i];
        for (int64_t k = 0; k < d; ++k) y[k] = a[k] * res;
        for (int64_t k = 0; k < d; ++k) y[k] *= lam;
        return (lam / 2) * (res * res);
    } else if (p->loss_kind == ORC_LOSS_LOGISTIC) {
        double yi = p->b[i], mu = p->lam[i];
        double u = orc_dot(a, x, d);
        double expyx = exp(yi * u);
        double c = -mu * yi / (1 + expyx);
        for (int64_t k = 0; k < d; ++k) y[k] = a[k] * c;
        return mu * log(1 + 1 / expyx);
    } else {
        const double *q = p->b + i * p->lda;
        double val = 0.0;
        for (int64_t k = 0; k < d; ++k) {
            double xk = x[k];
            double g1 = a[k] * xk + q[k];                 /* Quadratic: Qx + q  */
            double pr = xk < p->box_lo ? p->box_lo : (xk > p->box_hi ? p->box_hi : xk);
            double dist = xk - pr;
            double g2 = p->eta * dist;                    /* SqrDistL2          */
            y[k] = g1 + g2;
            val += 0.5 * a[k] * xk * xk + q[k] * xk + (p->eta / 2) * dist * dist;
        }
        return val;
    }
}

/* prox!(y, g, x, γ) */
void orc_prox(const orc_problem *p, double *y, const double *x, double gamma) {
    const int64_t d = p->d;
    if (p->reg_kind == ORC_REG_NORML1) {
        double gl = gamma * p->reg_lambda;
        for (int64_t k = 0; k < d; ++k) {
            double xk = x[k];
            y[k] = xk + (xk <= -gl ? gl : (xk >= gl ? -gl : -xk));
        }
    } else if (p->reg_kind == ORC_REG_NORML1_PAIRS) {
        /* complex NormL1 (normL1.jl, complex method): y_k = sign(x_k)·max(0, |x_k| − γλ), sign(z) = z/|z|, on (re, im) pairs */
        double gl = gamma * p->reg_lambda;
        for (int64_t k = 0; k + 1 < d; k += 2) {
            double re = x[k], im = x[k + 1];
            double ab = hypot(re, im);
            double m = ab - gl > 0 ? ab - gl : 0.0;
            y[k] = ab == 0 ? 0.0 : (re / ab) * m;
            y[k + 1] = ab == 0 ? 0.0 : (im / ab) * m;
        }
    } else if (p->reg_kind == ORC_REG_INDBOX) {
        for (int64_t k = 0; k < d; ++k) {
            double lo = p->reg_lo ? p->reg_lo[k] : p->reg_lo_s;
            double hi = p->reg_hi ? p->reg_hi[k] : p->reg_hi_s;
            double xk = x[k];
            y[k] = xk < lo ? lo : (xk > hi ? hi : xk);
        }
    } else {
        if (y != x) memcpy(y, x, (size_t)d * sizeof(double));
    }
}

double orc_reg_value(const orc_problem *p, const double *x) {
    if (p->reg_kind == ORC_REG_NORML1_PAIRS) {
        double s = 0.0;
        for (int64_t k = 0; k + 1 < p->d; k += 2) s += hypot(x[k], x[k + 1]);
        return p->reg_lambda * s;
    }
    if (p->reg_kind == ORC_REG_NORML1) {
        double s = 0.0;
        for (int64_t k = 0; k < p->d; ++k) s += fabs(x[k]);
        return p->reg_lambda * s;
    }
    return 0.0;
}

/* (1/N) Σ f_i(x) and g(x) — the cost the reference's tests evaluate
 * (test/test_lasso.jl:45). */
void orc_objective(const orc_problem *p, const double *x, double *f_mean, double *g_val) {
    double *tmp = (double *)malloc((size_t)p->d * sizeof(double));
    double s = 0.0;
    for (int64_t i = 0; i < p->N; ++i) s += orc_gradient(p, i, x, tmp);
    free(tmp);
    *f_mean = s / (double)p->N;
    *g_val = orc_reg_value(p, x);
}

/* out = scale · Σ_i ∇f_i(x), accumulated in the reference's order
 * (SVRG_basic.jl:58-63: each gradient scaled first, then added). */
void orc_full_gradient(const orc_problem *p, const double *x, double scale, double *out) {
    const int64_t d = p->d;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    memset(out, 0, (size_t)d * sizeof(double));
    for (int64_t i = 0; i < p->N; ++i) {
        orc_gradient(p, i, x, g);
        for (int64_t k = 0; k < d; ++k) g[k] *= scale;
        for (int64_t k = 0; k < d; ++k) out[k] += g[k];
    }
    free(g);
}

/* ------------------------------------------------------------------------ */
/* Multi-threaded full-gradient passes.  NOT a restatement of the reference  */
/* (which is single-threaded, SVRG_basic.jl:58-63): (a) the generous,        */
/* separately labelled all-cores CPU baseline SURVEY.md §8d allows, and (b)  */
/* the full-scale parity check of the CUDA pass (N = 2^22 × 4096 does not    */
/* fit host memory: rows are regenerated on the fly from the generator).     */
/* Plain pthreads (the image's default CC has no libgomp spec).              */
/* Rows are split into nthreads contiguous chunks; every thread runs the     */
/* reference's per-row operation sequence into its own accumulator; the      */
/* accumulators are added in thread order (deterministic for a given count). */
/* ------------------------------------------------------------------------ */
#include <pthread.h>
#include <unistd.h>

int orc_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}

typedef struct {
    void (*fn)(int t, int T, void *ctx);
    int t, T;
    void *ctx;
} orc_task;
static void *orc_task_main(void *arg) {
    orc_task *k = (orc_task *)arg;
    k->fn(k->t, k->T, k->ctx);
    return NULL;
}
/* runs fn(t, T, ctx) for t = 0 … T−1 on T threads (plain pthreads: the image's default compiler has no libgomp spec) */
static void orc_parallel(int T, void (*fn)(int, int, void *), void *ctx) {
    if (T < 1) T = 1;
    pthread_t *th = (pthread_t *)malloc((size_t)T * sizeof(pthread_t));
    orc_task *tk = (orc_task *)malloc((size_t)T * sizeof(orc_task));
    for (int t = 0; t < T; ++t) {
        tk[t].fn = fn; tk[t].t = t; tk[t].T = T; tk[t].ctx = ctx;
        if (t > 0 && pthread_create(&th[t], NULL, orc_task_main, &tk[t]) != 0) th[t] = 0, fn(t, T, ctx);
    }
    fn(0, T, ctx);
    for (int t = 1; t < T; ++t)
        if (th[t]) pthread_join(th[t], NULL);
    free(th);
    free(tk);
}

typedef struct {
    const orc_problem *p;
    const double *x;
    double scale;
    double *acc;
} orc_fg_ctx;
static void orc_fg_chunk(int t, int T, void *vc) {
    orc_fg_ctx *c = (orc_fg_ctx *)vc;
    const int64_t d = c->p->d, N = c->p->N;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    double *a = c->acc + (size_t)t * d;
    const int64_t lo = N * t / T, hi = N * (t + 1) / T;
    for (int64_t i = lo; i < hi; ++i) {
        orc_gradient(c->p, i, c->x, g);
        for (int64_t k = 0; k < d; ++k) g[k] *= c->scale;
        for (int64_t k = 0; k < d; ++k) a[k] += g[k];
    }
    free(g);
}
void orc_full_gradient_omp(const orc_problem *p, const double *x, double scale, double *out, int nthreads) {
    const int64_t d = p->d;
    if (nthreads < 1) nthreads = 1;
    double *acc = (double *)calloc((size_t)nthreads * (size_t)d, sizeof(double));
    orc_fg_ctx c = {p, x, scale, acc};
    orc_parallel(nthreads, orc_fg_chunk, &c);
    memset(out, 0, (size_t)d * sizeof(double));
    for (int t = 0; t < nthreads; ++t)
        for (int64_t k = 0; k < d; ++k) out[k] += acc[(size_t)t * d + k];
    free(acc);
}

/* out = scale · Σ_{i in [i0, i0+n)} ∇f_i(x) and *f_sum = Σ f_i(x) for the synthetic row problem (kind, d, seed) with
 * λ_i = μ_i = lam for all i, rows regenerated on the fly (never stored).  Long-double accumulation of the row dot and
 * of the per-thread sums, so that the result is a reference for the GPU pass rather than a peer. */
typedef struct {
    int kind;
    int64_t d, i0, n;
    uint64_t seed;
    double lam;
    const double *x;
    long double *acc;
} orc_syn_ctx;
static void orc_syn_chunk(int t, int T, void *vc) {
    orc_syn_ctx *c = (orc_syn_ctx *)vc;
    const int64_t d = c->d;
    const int kind = c->kind;
    double *row = (double *)malloc((size_t)d * sizeof(double));
    long double *a = c->acc + (size_t)t * (d + 1);
    const int64_t lo = c->i0 + c->n * t / T, hi = c->i0 + c->n * (t + 1) / T;
    for (int64_t i = lo; i < hi; ++i) {
        long double u = 0.0L;
        for (int64_t j = 0; j < d; ++j) {
            row[j] = ciao_syn_entry(kind, d, c->seed, i, j);
            u += (long double)row[j] * (long double)c->x[j];
        }
        const double rhs = ciao_syn_rhs(kind, d, c->seed, i);
        double cf;
        if (kind == CIAO_SYN_LASSO) {
            const double res = (double)u - rhs;
            cf = res * c->lam;
            a[d] += (long double)(c->lam / 2) * ((long double)res * res);
        } else {
            const double e = exp(rhs * (double)u);
            cf = -c->lam * rhs / (1 + e);
            a[d] += (long double)c->lam * logl(1 + 1 / (long double)e);
        }
        for (int64_t j = 0; j < d; ++j) a[j] += (long double)row[j] * cf;
    }
    free(row);
}
void orc_full_gradient_synth_omp(int kind, int64_t d, uint64_t seed, double lam, int64_t i0, int64_t n, const double *x,
                                 double scale, double *out, double *f_sum, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    long double *acc = (long double *)calloc((size_t)nthreads * (size_t)(d + 1), sizeof(long double));
    orc_syn_ctx c = {kind, d, i0, n, seed, lam, x, acc};
    orc_parallel(nthreads, orc_syn_chunk, &c);
    for (int64_t k = 0; k <= d; ++k) {
        long double s = 0.0L;
        for (int t = 0; t < nthreads; ++t) s += acc[(size_t)t * (d + 1) + k];
        if (k < d) out[k] = (double)(s * scale);
        else if (f_sum) *f_sum = (double)s;
    }
    free(acc);
}

/* Julia's sum over a Vector of d-vectors with stride ld (Base.mapreduce_impl:
 * pairwise, sequential below 1024 terms).  out = Σ_{i in [lo,hi)} v_i */
static void orc_pairwise_sum(const double *v, int64_t ld, int64_t d, int64_t lo, int64_t hi,
                             const double *div, double *out) {
    if (hi - lo < 1024) {
        for (int64_t k = 0; k < d; ++k) out[k] = div ? v[lo * ld + k] / div[lo] : v[lo * ld + k];
        for (int64_t i = lo + 1; i < hi; ++i)
            for (int64_t k = 0; k < d; ++k)
                out[k] += div ? v[i * ld + k] / div[i] : v[i * ld + k];
    } else {
        int64_t mid = lo + ((hi - lo) >> 1);
        double *r = (double *)malloc((size_t)d * sizeof(double));
        orc_pairwise_sum(v, ld, d, lo, mid, div, out);
        orc_pairwise_sum(v, ld, d, mid, hi, div, r);
        for (int64_t k = 0; k < d; ++k) out[k] += r[k];
        free(r);
    }
}

static double orc_pairwise_scalar(const double *v, int64_t lo, int64_t hi, int recip) {
    if (hi - lo < 1024) {
        double s = recip ? 1 / v[lo] : v[lo];
        for (int64_t i = lo + 1; i < hi; ++i) s += recip ? 1 / v[i] : v[i];
        return s;
    }
    int64_t mid = lo + ((hi - lo) >> 1);
    return orc_pairwise_scalar(v, lo, mid, recip) + orc_pairwise_scalar(v, mid, hi, recip);
}

/* γ̂ = 1/sum(1 ./ γ)  (Finito_basic.jl:82, Finito_LFinito.jl:66) */
double orc_finito_hat_gamma(const double *gamma, int64_t N) {
    return 1 / orc_pairwise_scalar(gamma, 0, N, 1);
}
/* γ̂ = sum(γ)  (ProShI_basic.jl:82) */
double orc_proshi_hat_gamma(const double *gamma, int64_t N) {
    return orc_pairwise_scalar(gamma, 0, N, 0);
}

/* ------------------------------------------------------------------------ */
/* SVRG / SVRG++   (SVRG_basic.jl)                                           */
/* ------------------------------------------------------------------------ */

/* SVRG_basic.jl:58-66 — av = Σ ∇f_i(x0)/N ; z_full = x0 ; z = 0 ; w = x0 */
void orc_svrg_init(const orc_problem *p, const double *x0, double *av, double *z,
                   double *z_full, double *w) {
    const int64_t d = p->d, N = p->N;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    memset(av, 0, (size_t)d * sizeof(double));
    for (int64_t i = 0; i < N; ++i) {
        orc_gradient(p, i, x0, g);
        for (int64_t k = 0; k < d; ++k) g[k] /= (double)N;
        for (int64_t k = 0; k < d; ++k) av[k] += g[k];
    }
    free(g);
    memcpy(z_full, x0, (size_t)d * sizeof(double));
    memset(z, 0, (size_t)d * sizeof(double));
    memcpy(w, x0, (size_t)d * sizeof(double));
}

/* SVRG_basic.jl:73-82 — the m sequential inner steps on the given indices */
void orc_svrg_inner(const orc_problem *p, double gamma, const int64_t *idx1, int64_t m,
                    const double *av, double *z, const double *z_full, double *w) {
    const int64_t d = p->d;
    double *temp = (double *)malloc((size_t)d * sizeof(double));
    double *gtmp = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t s = 0; s < m; ++s) {
        int64_t i = idx1[s] - 1;
        orc_gradient(p, i, z_full, temp);                          /* :74 */
        orc_gradient(p, i, w, gtmp);                               /* :75 */
        for (int64_t k = 0; k < d; ++k) temp[k] -= gtmp[k];        /* :76 */
        for (int64_t k = 0; k < d; ++k) temp[k] -= av[k];          /* :77 */
        for (int64_t k = 0; k < d; ++k) temp[k] *= gamma;          /* :78 */
        for (int64_t k = 0; k < d; ++k) temp[k] += w[k];           /* :79 */
        orc_prox(p, w, temp, gamma);                               /* :80 */
        for (int64_t k = 0; k < d; ++k) z[k] += w[k];              /* :81 */
    }
    free(temp);
    free(gtmp);
}

/* SVRG_basic.jl:71-96 — one outer iteration (inner loop + snapshot + full gradient).
 * The caller doubles m when plus (SVRG_basic.jl:93). */
void orc_svrg_epoch(const orc_problem *p, double gamma, int plus, const int64_t *idx1, int64_t m,
                    double *av, double *z, double *z_full, double *w) {
    const int64_t d = p->d, N = p->N;
    orc_svrg_inner(p, gamma, idx1, m, av, z, z_full, w);
    for (int64_t k = 0; k < d; ++k) z_full[k] = z[k] / (double)m;  /* :84 */
    if (!plus) memcpy(w, z_full, (size_t)d * sizeof(double));      /* :85 */
    memset(z, 0, (size_t)d * sizeof(double));                      /* :86 */
    memset(av, 0, (size_t)d * sizeof(double));                     /* :87 */
    double *g = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t i = 0; i < N; ++i) {                              /* :88-92 */
        orc_gradient(p, i, z_full, g);
        for (int64_t k = 0; k < d; ++k) g[k] /= (double)N;
        for (int64_t k = 0; k < d; ++k) av[k] += g[k];
    }
    free(g);
}

/* ------------------------------------------------------------------------ */
/* SAGA / SAG   (SAGA_basic.jl)                                              */
/* ------------------------------------------------------------------------ */

/* SAGA_basic.jl:41-48 — table s_i = ∇f_i(x0); av = sum(s)/N; z = prox_g((1−γ)x0, γ) */
void orc_saga_init(const orc_problem *p, const double *x0, double gamma, double *s, double *av,
                   double *z) {
    const int64_t d = p->d, N = p->N;
    for (int64_t i = 0; i < N; ++i) orc_gradient(p, i, x0, s + i * d);
    orc_pairwise_sum(s, d, d, 0, N, NULL, av);
    for (int64_t k = 0; k < d; ++k) av[k] /= (double)N;
    double *t = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t k = 0; k < d; ++k) t[k] = (1 - gamma) * x0[k];
    orc_prox(p, z, t, gamma);
    free(t);
}

/* SAGA_basic.jl:53-68 — K sequential single-sample steps */
void orc_saga_steps(const orc_problem *p, double gamma, int sag, const int64_t *idx1, int64_t K,
                    double *s, double *av, double *z) {
    const int64_t d = p->d;
    const double N = (double)p->N;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    double *w = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t t = 0; t < K; ++t) {
        int64_t i = idx1[t] - 1;
        double *si = s + i * d;
        orc_gradient(p, i, z, g);                                               /* :56 */
        if (sag) {
            for (int64_t k = 0; k < d; ++k) av[k] += (g[k] - si[k]) / N;        /* :58 */
            for (int64_t k = 0; k < d; ++k) w[k] = z[k] - gamma * av[k];        /* :59 */
        } else {
            for (int64_t k = 0; k < d; ++k) w[k] = z[k] - gamma * (g[k] - si[k] + av[k]); /* :61 */
            for (int64_t k = 0; k < d; ++k) av[k] += (g[k] - si[k]) / N;        /* :62 */
        }
        orc_prox(p, z, w, gamma);                                               /* :64 */
        memcpy(si, g, (size_t)d * sizeof(double));                              /* :65 */
    }
    free(g);
    free(w);
}

/* ------------------------------------------------------------------------ */
/* Finito / MISO / DIAG basic   (Finito_basic.jl)                            */
/* ------------------------------------------------------------------------ */

/* Finito_basic.jl:76-84 — s_i = x0 − (γ_i/N)∇f_i(x0); av = γ̂·sum(s ./ γ); z = prox_g(av, γ̂) */
void orc_finito_init(const orc_problem *p, const double *x0, const double *gamma, double hat_gamma,
                     double *s, double *av, double *z) {
    const int64_t d = p->d, N = p->N;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t i = 0; i < N; ++i) {
        orc_gradient(p, i, x0, g);
        double c = gamma[i] / (double)N;
        for (int64_t k = 0; k < d; ++k) s[i * d + k] = x0[k] - c * g[k];
    }
    free(g);
    orc_pairwise_sum(s, d, d, 0, N, gamma, av);
    for (int64_t k = 0; k < d; ++k) av[k] = hat_gamma * av[k];
    orc_prox(p, z, av, hat_gamma);
}

/* Finito_basic.jl:110-118 — batches in CSR form: batch j = idx1[ptr[j] .. ptr[j+1]) ;
 * the prox closes every batch. Index selection (:96-108) is the caller's. */
void orc_finito_steps(const orc_problem *p, const double *gamma, double hat_gamma,
                      const int64_t *idx1, const int64_t *ptr, int64_t n_batches, double *s,
                      double *av, double *z) {
    const int64_t d = p->d;
    const double N = (double)p->N;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t j = 0; j < n_batches; ++j) {
        for (int64_t t = ptr[j]; t < ptr[j + 1]; ++t) {
            int64_t i = idx1[t] - 1;
            double *si = s + i * d;
            orc_gradient(p, i, z, g);                                           /* :112 */
            double c = -(gamma[i] / N);
            for (int64_t k = 0; k < d; ++k) g[k] *= c;                          /* :113 */
            for (int64_t k = 0; k < d; ++k) g[k] += z[k];                       /* :114 */
            double r = hat_gamma / gamma[i];
            for (int64_t k = 0; k < d; ++k) av[k] += (g[k] - si[k]) * r;        /* :115 */
            memcpy(si, g, (size_t)d * sizeof(double));                          /* :116 */
        }
        orc_prox(p, z, av, hat_gamma);                                          /* :118 */
    }
    free(g);
}

/* ------------------------------------------------------------------------ */
/* Finito adaptive   (Finito_adaptive.jl)                                     */
/* ------------------------------------------------------------------------ */

/* LinearAlgebra.norm of a d-vector (2-norm; BLAS nrm2 in the reference, restated as sqrt of the left-to-right
 * sum of squares — the inputs here are far from over/underflow) */
static double orc_norm2(const double *v, int64_t d) {
    double s = 0.0;
    for (int64_t k = 0; k < d; ++k) s += v[k] * v[k];
    return sqrt(s);
}

/* Finito_adaptive.jl:59-99.  Tables: s (N×d, x_i), gf (N×d, ∇f_i(x_i)), fi_x (N), gamma (N, out).
 * Returns 0, or −1 when ∇f_i(x0 + 1) == ∇f_i(x0) for some i: the reference then draws random perturbations
 * (:75-81, global RNG) — unsupported here as in the engine. */
/* The random restart of the stepsize estimate (:77-83) draws from Julia's global RNG; the draw stays with the caller:
 * perturb(user, i (1-based), t, xeps) must fill xeps = x0 .+ rand(t * [-1, 1], size(x0)) and return 0.  NULL: a degenerate
 * component (∇f_i(x0 + 1) == ∇f_i(x0)) makes the init return -1. */
typedef int (*orc_perturb_fn)(void *user, int64_t i1, int64_t t, double *xeps);

int orc_finito_adaptive_init_cb(const orc_problem *p, const double *x0, double alpha, double *s, double *gf,
                                double *fi_x, double *gamma, double *hat_gamma, double *av, double *z,
                                orc_perturb_fn perturb, void *user) {
    const int64_t d = p->d, N = p->N;
    double *xeps = (double *)malloc((size_t)d * sizeof(double));
    double *ge = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t i = 0; i < N; ++i) {                                            /* :65-68 */
        fi_x[i] = orc_gradient(p, i, x0, gf + i * d);
        memcpy(s + i * d, x0, (size_t)d * sizeof(double));
    }
    for (int64_t i = 0; i < N; ++i) {                                            /* :71-87 */
        for (int64_t k = 0; k < d; ++k) xeps[k] = x0[k] + 1.0;                   /* :73 */
        orc_gradient(p, i, xeps, ge);
        for (int64_t k = 0; k < d; ++k) ge[k] -= gf[i * d + k];
        double nmg = orc_norm2(ge, d);                                           /* :75 */
        int64_t t = 1;                                                           /* :76 */
        while (nmg < 2.220446049250313e-16) {                                    /* :77 eps(R) */
            if (!perturb || perturb(user, i + 1, t, xeps) != 0) { free(xeps); free(ge); return -1; }   /* :79 */
            orc_gradient(p, i, xeps, ge);                                        /* :80 */
            for (int64_t k = 0; k < d; ++k) ge[k] -= gf[i * d + k];
            nmg = orc_norm2(ge, d);                                              /* :81 */
            t *= 2;                                                              /* :82 */
        }
        double L_int = nmg / ((double)t * sqrt((double)d));                      /* :84 */
        L_int /= (double)N;                                                      /* :85 */
        gamma[i] = alpha / L_int;                                                /* :86 */
    }
    *hat_gamma = 1 / orc_pairwise_scalar(gamma, 0, N, 1);                        /* :89 */
    double *sg = (double *)malloc((size_t)d * sizeof(double));
    orc_pairwise_sum(s, d, d, 0, N, gamma, av);                                  /* sum(s ./ γ) */
    orc_pairwise_sum(gf, d, d, 0, N, NULL, sg);                                  /* sum(∇f) */
    for (int64_t k = 0; k < d; ++k) av[k] = *hat_gamma * (av[k] - sg[k] / (double)N);  /* :90 */
    orc_prox(p, z, av, *hat_gamma);                                              /* :91 */
    free(xeps); free(ge); free(sg);
    return 0;
}

int orc_finito_adaptive_init(const orc_problem *p, const double *x0, double alpha, double *s, double *gf,
                             double *fi_x, double *gamma, double *hat_gamma, double *av, double *z) {
    return orc_finito_adaptive_init_cb(p, x0, alpha, s, gf, fi_x, gamma, hat_gamma, av, z, NULL, NULL);
}

/* Finito_adaptive.jl:101-160, K steps on the given (1-based) indices (selection :107-119 is the caller's).
 * Returns the number of steps completed: < K when γ_i fell below tol_b/N (:125-128, `return nothing`). */
int64_t orc_finito_adaptive_steps(const orc_problem *p, double alpha, double tol_b, const int64_t *idx1, int64_t K,
                                  double *s, double *gf, double *fi_x, double *gamma, double *hat_gamma,
                                  double *av, double *z, int64_t *n_backtracks) {
    const int64_t d = p->d;
    const double N = (double)p->N;
    double *res = (double *)malloc((size_t)d * sizeof(double));
    double *tmp = (double *)malloc((size_t)d * sizeof(double));
    double hg = *hat_gamma;
    int64_t done = 0, nbt = 0;
    for (int64_t t = 0; t < K; ++t) {
        const int64_t i = idx1[t] - 1;
        double *si = s + i * d, *gi = gf + i * d;
        for (int64_t k = 0; k < d; ++k) res[k] = z[k] - si[k];                   /* :121 */
        int stop = 0;
        for (;;) {                                                               /* :123-147 */
            if (gamma[i] < tol_b / N) { stop = 1; break; }                       /* :124-127 */
            double fi_z = orc_gradient(p, i, z, tmp);                            /* :128 (value only) */
            double nr = orc_norm2(res, d);
            double fi_model = fi_x[i] + orc_dot(gi, res, d) + (0.5 * N * alpha / gamma[i]) * (nr * nr);  /* :129-132 */
            double tol = 10 * 2.220446049250313e-16 * (1 + fabs(fi_z));          /* :133 */
            if (fi_z <= fi_model + tol) break;                                   /* :134 */
            double gamma_b = gamma[i];                                           /* :136 */
            gamma[i] *= 0.8;                                                     /* :137 */
            for (int64_t k = 0; k < d; ++k) av[k] /= hg;                         /* :139 */
            for (int64_t k = 0; k < d; ++k) av[k] += si[k] / gamma[i];           /* :140 */
            for (int64_t k = 0; k < d; ++k) av[k] -= si[k] / gamma_b;            /* :141 */
            hg = 1 / (1 / hg + 1 / gamma[i] - 1 / gamma_b);                      /* :142 */
            for (int64_t k = 0; k < d; ++k) av[k] *= hg;                         /* :143 */
            orc_prox(p, z, av, hg);                                              /* :144 */
            for (int64_t k = 0; k < d; ++k) res[k] = z[k] - si[k];               /* :145 */
            ++nbt;
        }
        if (stop) break;
        double r = hg / gamma[i];
        for (int64_t k = 0; k < d; ++k) av[k] += r * (z[k] - si[k]);             /* :149 */
        memcpy(si, z, (size_t)d * sizeof(double));                               /* :150 */
        double c = hg / N;
        for (int64_t k = 0; k < d; ++k) av[k] += c * gi[k];                      /* :151 */
        fi_x[i] = orc_gradient(p, i, z, gi);                                     /* :152 */
        for (int64_t k = 0; k < d; ++k) av[k] -= c * gi[k];                      /* :153 */
        orc_prox(p, z, av, hg);                                                  /* :154 */
        ++done;
    }
    *hat_gamma = hg;
    if (n_backtracks) *n_backtracks = nbt;
    free(res); free(tmp);
    return done;
}

/* ------------------------------------------------------------------------ */
/* LFinito   (Finito_LFinito.jl)                                             */
/* ------------------------------------------------------------------------ */

/* Finito_LFinito.jl:67-72 — av = x0 − Σ (γ̂/N)∇f_i(x0); z = z_full = copy(av) (ctor :33-35) */
void orc_lfinito_init(const orc_problem *p, const double *x0, double hat_gamma, double *av,
                      double *z, double *z_full) {
    const int64_t d = p->d, N = p->N;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    memcpy(av, x0, (size_t)d * sizeof(double));
    double c = hat_gamma / (double)N;
    for (int64_t i = 0; i < N; ++i) {
        orc_gradient(p, i, x0, g);
        for (int64_t k = 0; k < d; ++k) g[k] *= c;
        for (int64_t k = 0; k < d; ++k) av[k] -= g[k];
    }
    free(g);
    memcpy(z, av, (size_t)d * sizeof(double));
    memcpy(z_full, av, (size_t)d * sizeof(double));
}

/* Finito_LFinito.jl:78-103 — one outer iteration.  batch_order1 = state.inds
 * (1-based batch numbers, after the optional randperm :89); batch j covers rows
 * r(j−1)+1 .. min(jr, N) (:44-49). */
void orc_lfinito_outer(const orc_problem *p, const double *gamma, double hat_gamma,
                       const int64_t *batch_order1, int64_t n_batches, int64_t r, double *av,
                       double *z, double *z_full) {
    const int64_t d = p->d, N = p->N;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    double c = hat_gamma / (double)N;
    orc_prox(p, z_full, av, hat_gamma);                                         /* :83 */
    memcpy(av, z_full, (size_t)d * sizeof(double));                             /* :84 */
    for (int64_t i = 0; i < N; ++i) {                                           /* :85-88 */
        orc_gradient(p, i, z_full, g);
        for (int64_t k = 0; k < d; ++k) av[k] -= c * g[k];
    }
    for (int64_t jj = 0; jj < n_batches; ++jj) {                                /* :91 */
        int64_t j = batch_order1[jj] - 1;
        orc_prox(p, z, av, hat_gamma);                                          /* :92 */
        int64_t lo = r * j, hi = lo + r < N ? lo + r : N;
        for (int64_t i = lo; i < hi; ++i) {
            orc_gradient(p, i, z_full, g);                                      /* :94 */
            for (int64_t k = 0; k < d; ++k) av[k] += c * g[k];                  /* :95 */
            orc_gradient(p, i, z, g);                                           /* :96 */
            for (int64_t k = 0; k < d; ++k) av[k] -= c * g[k];                  /* :97 */
            double rr = hat_gamma / gamma[i];
            for (int64_t k = 0; k < d; ++k) av[k] += rr * (z[k] - z_full[k]);   /* :98 */
        }
    }
    free(g);
}

/* ------------------------------------------------------------------------ */
/* ProShI   (ProShI_basic.jl)                                                */
/* ------------------------------------------------------------------------ */

static void orc_proshi_dual(const orc_problem *p, double hat_gamma, const double *av, double *z) {
    const int64_t d = p->d;
    orc_prox(p, z, av, hat_gamma);                         /* :84 / :121 */
    for (int64_t k = 0; k < d; ++k) z[k] -= av[k];         /* :85 / :122 */
    for (int64_t k = 0; k < d; ++k) z[k] /= hat_gamma;     /* :86 / :123 */
}

/* ProShI_basic.jl:76-86 — s_i = x0 − (γ_i/N)∇f_i(x0); av = sum(s); z = (prox_g(av,γ̂) − av)/γ̂ */
void orc_proshi_init(const orc_problem *p, const double *x0, const double *gamma, double hat_gamma,
                     double *s, double *av, double *z) {
    const int64_t d = p->d, N = p->N;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t i = 0; i < N; ++i) {
        orc_gradient(p, i, x0, g);
        double c = gamma[i] / (double)N;
        for (int64_t k = 0; k < d; ++k) s[i * d + k] = x0[k] - c * g[k];
    }
    free(g);
    orc_pairwise_sum(s, d, d, 0, N, NULL, av);
    orc_proshi_dual(p, hat_gamma, av, z);
}

/* ProShI_basic.jl:111-123 — CSR batches as in orc_finito_steps */
void orc_proshi_steps(const orc_problem *p, const double *gamma, double hat_gamma,
                      const int64_t *idx1, const int64_t *ptr, int64_t n_batches, double *s,
                      double *av, double *z) {
    const int64_t d = p->d;
    const double N = (double)p->N;
    double *g = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t j = 0; j < n_batches; ++j) {
        for (int64_t t = ptr[j]; t < ptr[j + 1]; ++t) {
            int64_t i = idx1[t] - 1;
            double *si = s + i * d;
            for (int64_t k = 0; k < d; ++k) av[k] -= si[k];                     /* :113 */
            for (int64_t k = 0; k < d; ++k) si[k] += gamma[i] * z[k];           /* :114 */
            orc_gradient(p, i, si, g);                                          /* :115 */
            double c = -(gamma[i] / N);
            for (int64_t k = 0; k < d; ++k) g[k] *= c;                          /* :116 */
            for (int64_t k = 0; k < d; ++k) g[k] += si[k];                      /* :117 */
            for (int64_t k = 0; k < d; ++k) av[k] += g[k];                      /* :118 */
            memcpy(si, g, (size_t)d * sizeof(double));                          /* :119 */
        }
        orc_proshi_dual(p, hat_gamma, av, z);                                   /* :121-123 */
    }
    free(g);
}

/* ProShI_basic.jl:127-132 — IN-PLACE s_i += γ_i z for all i, every call */
void orc_proshi_solution(const orc_problem *p, const double *gamma, const double *z, double *s) {
    const int64_t d = p->d;
    for (int64_t i = 0; i < p->N; ++i)
        for (int64_t k = 0; k < d; ++k) s[i * d + k] += gamma[i] * z[k];
}

/* ------------------------------------------------------------------------ */
/* synthetic inputs (bit-identical to the device generator, include/ciao_gen.h) */
/* ------------------------------------------------------------------------ */

typedef struct {
    int kind;
    int64_t d, i0, n;
    uint64_t seed;
    double *A, *rhs;
} orc_gen_ctx;
static void orc_gen_chunk(int t, int T, void *vc) {
    orc_gen_ctx *c = (orc_gen_ctx *)vc;
    for (int64_t r = c->n * t / T; r < c->n * (t + 1) / T; ++r) {
        int64_t i = c->i0 + r;
        for (int64_t j = 0; j < c->d; ++j) c->A[r * c->d + j] = ciao_syn_entry(c->kind, c->d, c->seed, i, j);
        if (c->rhs && c->kind != CIAO_SYN_SHARING) c->rhs[r] = ciao_syn_rhs(c->kind, c->d, c->seed, i);
    }
}
/* every entry is a pure function of (seed, i, j): the same bits with any thread count */
void orc_gen_rows(int kind, int64_t d, uint64_t seed, int64_t i0, int64_t n, double *A, double *rhs) {
    orc_gen_ctx c = {kind, d, i0, n, seed, A, rhs};
    orc_parallel(n * d >= (1 << 22) ? orc_num_threads() : 1, orc_gen_chunk, &c);
}

void orc_gen_xtrue(int kind, int64_t d, uint64_t seed, double *x) {
    memset(x, 0, (size_t)d * sizeof(double));
    int64_t p = ciao_syn_support_size(kind, d);
    for (int64_t t = 0; t < p; ++t)
        x[ciao_syn_support_pos(kind, d, seed, t)] = ciao_syn_support_val(seed, t);
}

/* max_i ‖a_i‖² over rows [0,N) — L_i = λ_i‖a_i‖² (LS, test_lasso.jl:55) or
 * 0.25‖a_i‖² (logistic, test_logistic_l1.jl:39) */
double orc_max_row_sqnorm(const orc_problem *p) {
    double mx = 0.0;
    for (int64_t i = 0; i < p->N * (p->M > 1 ? p->M : 1); ++i) {
        const double *a = p->A + i * p->lda;
        double s = orc_dot(a, a, p->d);
        if (s > mx) mx = s;
    }
    return mx;
}

/* ciao_gen.h — counter-based synthetic problem generator shared by the CUDA
 * engine (device + host code in libciao_cuda) and by the CPU oracle.
 *
 * Every value is a pure function of (seed, index), so any GPU shard, the host
 * and the oracle produce bit-identical inputs without communicating
 * (SURVEY.md §8d "Concrete synthetic inputs").  The three problem shapes scale
 * up the reference's own test problems:
 *   CIAO_SYN_LASSO    — test/test_lasso.jl:15-60   (LeastSquares rows + NormL1)
 *   CIAO_SYN_LOGISTIC — test/test_logistic_l1.jl:12-46 (Logistic rows, bias column)
 *   CIAO_SYN_SHARING  — test/test_sharing.jl:9-26  (diag Quadratic + SqrDistL2 blocks)
 *
 * Only +,*,fma on exactly representable operands in a fixed order are used, so
 * host and device agree bit-for-bit (fma() is correctly rounded on both).
 */
#ifndef CIAO_GEN_H
#define CIAO_GEN_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define CIAO_HD __host__ __device__ __forceinline__
#else
#define CIAO_HD static inline
#endif

#define CIAO_SYN_LASSO 0
#define CIAO_SYN_LOGISTIC 1
#define CIAO_SYN_SHARING 2

/* number of planted non-zeros in x_true (Lasso: p, logistic: 5 % of d, min 1) */
#define CIAO_SYN_LASSO_P 64

/* (k+1)-th output of a splitmix64 stream seeded with `seed` */
CIAO_HD uint64_t ciao_hash64(uint64_t seed, uint64_t k) {
    uint64_t z = seed + (k + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* uniform in [0,1) with 53 random bits */
CIAO_HD double ciao_u01(uint64_t seed, uint64_t k) {
    return (double)(ciao_hash64(seed, k) >> 11) * (1.0 / 9007199254740992.0);
}

/* sub-stream seeds */
#define CIAO_SEED_A(s) ((s) ^ 0xA5A5A5A5A5A5A5A5ull)
#define CIAO_SEED_X(s) ((s) ^ 0x0123456789ABCDEFull)
#define CIAO_SEED_B(s) ((s) ^ 0x5DEECE66D1CE4E5Bull)

CIAO_HD int64_t ciao_syn_support_size(int kind, int64_t d) {
    if (kind == CIAO_SYN_LASSO) return d < CIAO_SYN_LASSO_P ? d : CIAO_SYN_LASSO_P;
    int64_t p = d / 20;
    return p < 1 ? 1 : p;
}

/* t-th support position of x_true: one position inside each of p equal strata */
CIAO_HD int64_t ciao_syn_support_pos(int kind, int64_t d, uint64_t seed, int64_t t) {
    int64_t p = ciao_syn_support_size(kind, d);
    int64_t stride = d / p;
    return t * stride + (int64_t)(ciao_hash64(CIAO_SEED_X(seed), (uint64_t)t) % (uint64_t)stride);
}

/* t-th non-zero value of x_true, in ±[0.5, 1.5) */
CIAO_HD double ciao_syn_support_val(uint64_t seed, int64_t t) {
    double mag = 0.5 + ciao_u01(CIAO_SEED_X(seed), 1000003ull + (uint64_t)t);
    return (ciao_hash64(CIAO_SEED_X(seed), 2000003ull + (uint64_t)t) & 1ull) ? mag : -mag;
}

/* entry (i,j) of the N×d data matrix.
 *   LASSO:    uniform in [-1,1)
 *   LOGISTIC: uniform in [-1,1) scaled by 2^-5 (≈ 1/√d at d=1024), last column ≡ 1 (bias)
 *   SHARING:  diagonal of Q_i: 10·u, with ≈1 % entries in (-1,0]            */
CIAO_HD double ciao_syn_entry(int kind, int64_t d, uint64_t seed, int64_t i, int64_t j) {
    uint64_t k = (uint64_t)i * (uint64_t)d + (uint64_t)j;
    double u = ciao_u01(CIAO_SEED_A(seed), k);
    if (kind == CIAO_SYN_LASSO) return 2.0 * u - 1.0;
    if (kind == CIAO_SYN_LOGISTIC) return (j == d - 1) ? 1.0 : (2.0 * u - 1.0) * 0.03125;
    /* sharing */
    uint64_t h = ciao_hash64(CIAO_SEED_B(seed), k);
    return ((h % 100ull) == 0ull) ? -u : 10.0 * u;
}

/* right-hand side b_i (Lasso: a_i·x_true + 0.01·noise) or label y_i ∈ {-1,+1}
 * (logistic: sign(a_i·x_true + 0.1·noise)).  The dot runs over the support only,
 * in order t = 0..p-1, with fma — identical on host and device.            */
CIAO_HD double ciao_syn_rhs(int kind, int64_t d, uint64_t seed, int64_t i) {
    int64_t p = ciao_syn_support_size(kind, d);
    double acc = 0.0;
    for (int64_t t = 0; t < p; ++t) {
        int64_t j = ciao_syn_support_pos(kind, d, seed, t);
        acc = fma(ciao_syn_entry(kind, d, seed, i, j), ciao_syn_support_val(seed, t), acc);
    }
    double noise = 2.0 * ciao_u01(CIAO_SEED_B(seed), 4000037ull + (uint64_t)i) - 1.0;
    if (kind == CIAO_SYN_LASSO) return fma(0.01, noise, acc);
    double s = fma(0.1, noise, acc);
    return s >= 0.0 ? 1.0 : -1.0;
}

#endif /* CIAO_GEN_H */

// jlrng.cu — host-side restatement of the index draws the reference makes with Julia's default RNG (SURVEY.md §8a row a19,
// Appendix C), for hosts without Julia: the Python twin of the shim (sampling.JuliaRNG) and the tests.  Pure host code, no CUDA.
//
//   rand(state.ind, m)                    SVRG_basic.jl:73        → ciao_jlrng_rand_range (m draws)
//   rand(1:N)                             SAGA_basic.jl:55, Finito_adaptive.jl:108
//   sample(1:N, k, replace=false)         Finito_basic.jl:97, ProShI_basic.jl:98      → ciao_jlrng_sample_norep
//   randperm(n)                           Finito_basic.jl:102, Finito_LFinito.jl:89, ProShI_basic.jl:103 → ciao_jlrng_randperm
//
// What is restated: Julia ≥ 1.7's default generator (task-local Xoshiro256++; Random/src/Xoshiro.jl) and, on top of its
// UInt64 stream, Random's range sampler SamplerRangeNDL (Lemire's nearly-divisionless method, the default for 64-bit integer
// ranges since Julia 1.5; Random/src/generation.jl), randperm! with its masked-rejection rand_lt (Random/src/misc.jl) and
// StatsBase 0.33's sample!(…; replace=false) dispatch (k = 1 → one draw, k = 2 → samplepair, n < 24k → Fisher–Yates, else
// self-avoiding; StatsBase/src/sampling.jl).  Seeding (Random.seed!(n): SHA-256 of the seed's UInt32 words → 4 state words) is
// done by the caller (Python: hashlib) and handed in as the four state words.
//
// PINNING.  The generator + seeding + Float64 conversion reproduce the known answer printed in Julia's own documentation of
// Random.seed! / Xoshiro (`Random.seed!(1234); rand(2)` → 0.32597672886359486, 0.5490511363155669): tests/test_host_logic.py.
// The integer samplers are restated from the sources cited above and could NOT be checked against a running Julia (none in
// this image): they are "unverified" until the Julia shim runs.  Julia ≤ 1.6 (MersenneTwister) draws a different stream; the
// reference's Project.toml allows both, so "the reference's RNG" is only defined relative to the Julia actually running it —
// which is why index sequences are an INPUT at the C ABI.
#include <stdint.h>

#include <unordered_set>
#include <vector>

#include "../../include/ciao_cuda.h"

namespace {
struct Xo {
    uint64_t s0, s1, s2, s3;
};
inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
// xoshiro256++ (Xoshiro.jl `rand(rng, UInt64)`)
inline uint64_t next(Xo &r) {
    const uint64_t res = rotl(r.s0 + r.s3, 23) + r.s0;
    const uint64_t t = r.s1 << 17;
    r.s2 ^= r.s0;
    r.s3 ^= r.s1;
    r.s1 ^= r.s2;
    r.s0 ^= r.s3;
    r.s2 ^= t;
    r.s3 = rotl(r.s3, 45);
    return res;
}
// rand(rng, SamplerRangeNDL(1:n)) − 1: uniform in [0, n), n ≥ 1 (generation.jl; s = range length as UInt64)
inline uint64_t rand_below(Xo &r, uint64_t s) {
    unsigned __int128 m = (unsigned __int128)next(r) * s;
    uint64_t l = (uint64_t)m;
    if (l < s) {
        const uint64_t t = (0 - s) % s;   // mod(-s, s) on UInt64
        while (l < t) {
            m = (unsigned __int128)next(r) * s;
            l = (uint64_t)m;
        }
    }
    return (uint64_t)(m >> 64);
}
// rand_lt(r, n, mask): (rand(r, UInt52Raw()) % Int) & mask until < n; UInt52Raw on Xoshiro = rand(UInt64) >>> 12
inline int64_t rand_lt(Xo &r, int64_t n, int64_t mask) {
    for (;;) {
        const int64_t x = (int64_t)(next(r) >> 12) & mask;
        if (x < n) return x;
    }
}
inline Xo load(const uint64_t *st) { return Xo{st[0], st[1], st[2], st[3]}; }
inline void store(uint64_t *st, const Xo &r) { st[0] = r.s0; st[1] = r.s1; st[2] = r.s2; st[3] = r.s3; }
}  // namespace

// next UInt64 outputs (the raw stream; Float64: (u >> 11)·2^-53)
extern "C" int ciao_jlrng_next_u64(uint64_t *state4, uint64_t *out, int64_t n) {
    if (!state4 || (!out && n > 0) || n < 0) return CIAO_ERR_INVALID;
    Xo r = load(state4);
    for (int64_t i = 0; i < n; ++i) out[i] = next(r);
    store(state4, r);
    return CIAO_OK;
}

// m draws of rand(1:N) (1-based), in order — also rand(collect(1:N), m)
extern "C" int ciao_jlrng_rand_range(uint64_t *state4, int64_t N, int64_t *out, int64_t m) {
    if (!state4 || (!out && m > 0) || N < 1 || m < 0) return CIAO_ERR_INVALID;
    Xo r = load(state4);
    for (int64_t i = 0; i < m; ++i) out[i] = (int64_t)rand_below(r, (uint64_t)N) + 1;
    store(state4, r);
    return CIAO_OK;
}

// randperm(n) (1-based)
extern "C" int ciao_jlrng_randperm(uint64_t *state4, int64_t n, int64_t *out) {
    if (!state4 || (!out && n > 0) || n < 0) return CIAO_ERR_INVALID;
    if (n == 0) return CIAO_OK;
    Xo r = load(state4);
    out[0] = 1;
    int64_t mask = 3;
    for (int64_t i = 2; i <= n; ++i) {
        const int64_t j = 1 + rand_lt(r, i, mask);
        if (i != j) out[i - 1] = out[j - 1];
        out[j - 1] = i;
        if (i == 1 + mask) mask = 2 * mask + 1;
    }
    store(state4, r);
    return CIAO_OK;
}

// StatsBase.sample(1:N, k, replace=false) (unordered), 1-based
extern "C" int ciao_jlrng_sample_norep(uint64_t *state4, int64_t N, int64_t k, int64_t *out) {
    if (!state4 || (!out && k > 0) || N < 1 || k < 0 || k > N) return CIAO_ERR_INVALID;
    Xo r = load(state4);
    if (k == 1) {
        out[0] = (int64_t)rand_below(r, (uint64_t)N) + 1;
    } else if (k == 2) {   // samplepair
        const int64_t i1 = (int64_t)rand_below(r, (uint64_t)N) + 1;
        const int64_t i2 = (int64_t)rand_below(r, (uint64_t)(N - 1)) + 1;
        out[0] = i1;
        out[1] = i2 == i1 ? N : i2;
    } else if (k > 0 && N < 24 * k) {   // fisher_yates_sample!
        std::vector<int64_t> inds((size_t)N);
        for (int64_t i = 0; i < N; ++i) inds[(size_t)i] = i + 1;
        for (int64_t i = 1; i <= k; ++i) {
            const int64_t j = i + (int64_t)rand_below(r, (uint64_t)(N - i + 1));   // rand(rng, i:n)
            const int64_t t = inds[(size_t)j - 1];
            inds[(size_t)j - 1] = inds[(size_t)i - 1];
            inds[(size_t)i - 1] = t;
            out[i - 1] = t;
        }
    } else if (k > 0) {   // self_avoid_sample!
        std::unordered_set<int64_t> seen;
        seen.reserve((size_t)k * 2);
        for (int64_t i = 0; i < k; ++i) {
            int64_t idx = (int64_t)rand_below(r, (uint64_t)N) + 1;
            while (seen.count(idx)) idx = (int64_t)rand_below(r, (uint64_t)N) + 1;
            out[i] = idx;
            seen.insert(idx);
        }
    }
    store(state4, r);
    return CIAO_OK;
}

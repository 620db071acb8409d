"""Host-side index selection — the RNG call sites of the reference stay on the host.

The engine takes index sequences as INPUTS at the C ABI (1-based int64, as
Julia produces them), so the sample path is whatever the host RNG draws:
  * the Julia shim (julia/CIAOAlgorithmsCUDA.jl) calls the reference's own
    ``rand`` / ``StatsBase.sample`` / ``randperm`` in the reference's order, so
    its index stream is the reference's bit for bit;
  * this Python mirror draws from ``numpy.random.Generator`` through the same
    call structure (one call per reference call site).

Reference call sites:
  rand(state.ind, state.m)              SVRG_basic.jl:73
  rand(1:iter.N)                        SAGA_basic.jl:55
  sample(1:N, batch, replace=false)     Finito_basic.jl:97, ProShI_basic.jl:98
  randperm(d)                           Finito_basic.jl:102, Finito_LFinito.jl:89, ProShI_basic.jl:103
  rand(1:N), randperm(N)                Finito_adaptive.jl:108, 113
"""
from __future__ import annotations

import numpy as np


class HostRNG:
    """The four draws the reference makes, 1-based like Julia."""

    def __init__(self, seed=0):
        self.g = np.random.default_rng(seed)

    def rand_range(self, N: int) -> int:                       # rand(1:N)
        return int(self.g.integers(1, N + 1))

    def rand_vec(self, N: int, m: int) -> np.ndarray:          # rand(collect(1:N), m)
        return self.g.integers(1, N + 1, size=m, dtype=np.int64)

    def sample_norep(self, N: int, k: int) -> np.ndarray:      # sample(1:N, k, replace=false)
        return (self.g.choice(N, size=k, replace=False) + 1).astype(np.int64)

    def randperm(self, n: int) -> np.ndarray:                  # randperm(n)
        return (self.g.permutation(n) + 1).astype(np.int64)

    def rand_pm(self, t: int, d: int) -> np.ndarray:           # rand(t * [-1, 1], d)   Finito_adaptive.jl:79
        """d draws from the two-element array [-t, t]: element k is v[rand(1:2)], one draw after the other."""
        return np.where(self.rand_vec(2, d) == 1, -float(t), float(t))


class JuliaRNG(HostRNG):
    """The same four draws from a restatement of Julia ≥ 1.7's default RNG (Xoshiro256++ seeded like ``Random.seed!(seed)``)
    and of Random's / StatsBase 0.33's samplers (csrc/jlrng.cu): with it the Python twin walks the index stream the Julia
    shim would draw for the same seed.  Pinned: the raw stream (known answer in Julia's docs, tests/test_host_logic.py);
    unverified against a running Julia: the integer samplers built on it.  Julia ≤ 1.6 (MersenneTwister) draws differently."""

    def __init__(self, seed=0):
        import hashlib
        import struct
        from . import _lib as L
        self.lib = L.load()
        words, n = [], int(seed)
        if n < 0:
            raise ValueError("Random.seed! takes a non-negative integer")
        while True:                                    # Random.make_seed(n): the seed's UInt32 words, little end first
            words.append(n & 0xFFFFFFFF)
            n >>= 32
            if n == 0:
                break
        digest = hashlib.sha256(b"".join(struct.pack("<I", w) for w in words)).digest()   # seed!(::Xoshiro, ::Vector{UInt32})
        self.state = np.array(struct.unpack("<4Q", digest), dtype=np.uint64)

    def _call(self, fn, *args):
        from ._lib import check
        check(fn(self.state.ctypes.data, *args))

    def next_u64(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        self._call(self.lib.ciao_jlrng_next_u64, out.ctypes.data, n)
        return out

    def rand_float(self, n: int) -> np.ndarray:               # rand(n) of Float64 one at a time: (u >>> 11)·2^-53
        return (self.next_u64(n) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53

    def rand_range(self, N: int) -> int:
        return int(self.rand_vec(N, 1)[0])

    def rand_vec(self, N: int, m: int) -> np.ndarray:
        out = np.empty(m, dtype=np.int64)
        self._call(self.lib.ciao_jlrng_rand_range, int(N), out.ctypes.data, int(m))
        return out

    def sample_norep(self, N: int, k: int) -> np.ndarray:
        out = np.empty(k, dtype=np.int64)
        self._call(self.lib.ciao_jlrng_sample_norep, int(N), int(k), out.ctypes.data)
        return out

    def randperm(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int64)
        self._call(self.lib.ciao_jlrng_randperm, int(n), out.ctypes.data)
        return out


def static_batches(N: int, r: int):
    """Finito_basic.jl:52-57 / Finito_LFinito.jl:44-49 / ProShI_basic.jl:52-57:
    batch j (1-based) = r(j−1)+1 .. jr, plus one remainder batch."""
    nb_full = N // r
    out = [np.arange(r * j + 1, r * (j + 1) + 1, dtype=np.int64) for j in range(nb_full)]
    if r * nb_full < N:
        out.append(np.arange(r * nb_full + 1, N + 1, dtype=np.int64))
    return out


def n_batches(N: int, r: int) -> int:
    return -(-N // r)  # cld(N, r)   Finito_basic.jl:59


class BatchSweeper:
    """Index selection of Finito_basic.jl:96-108 and ProShI_basic.jl:97-109.

    sweeping 1: a fresh ``sample(1:N, r, replace=false)`` every step;
    sweeping 2: cyclic over the static batches, ``idxr = mod(idxr, d) + 1`` from
                ``idxr = 1`` — the first batch visited is #2 (a reference quirk);
    sweeping 3: natural order 1..d for the first pass (``inds = 1:d``, ``idx = 0``),
                a fresh ``randperm(d)`` at the start of every later pass.
    """

    def __init__(self, N: int, batch: int, sweeping: int, rng: HostRNG):
        assert sweeping in (1, 2, 3)
        self.N, self.r, self.sweeping, self.rng = N, batch, sweeping, rng
        self.d = n_batches(N, batch)
        self.idxr, self.idx = 1, 0                      # Finito_basic.jl:38-39
        self.inds = np.arange(1, self.d + 1, dtype=np.int64)  # :40

    def batch_rows(self, j1: int) -> np.ndarray:
        lo = self.r * (j1 - 1) + 1
        hi = min(self.r * j1, self.N)
        return np.arange(lo, hi + 1, dtype=np.int64)

    def next(self) -> np.ndarray:
        if self.sweeping == 1:
            return self.rng.sample_norep(self.N, self.r)
        if self.sweeping == 2:
            self.idxr = self.idxr % self.d + 1
        else:
            if self.idx == self.d:
                self.inds = self.rng.randperm(self.d)
                self.idx = 1
            else:
                self.idx += 1
            self.idxr = int(self.inds[self.idx - 1])
        return self.batch_rows(self.idxr)

    def take(self, k: int):
        return [self.next() for _ in range(k)]


class AdaptiveSweeper:
    """Index selection of Finito_adaptive.jl:107-119 (single indices, no batches).

    sweeping 1: ``rand(1:N)``;  sweeping 2: ``idxr = mod(idxr, N) + 1`` from ``idxr = 0`` — starts at 1 (unlike the
    basic variant);  sweeping 3: natural order for the first pass (``ind = 1:N``, ``idx = 0``), then ``randperm(N)``.
    """

    def __init__(self, N: int, sweeping: int, rng: HostRNG):
        assert sweeping in (1, 2, 3)
        self.N, self.sweeping, self.rng = N, sweeping, rng
        self.idxr, self.idx = 0, 0                       # Finito_adaptive.jl:54-55
        self.ind = np.arange(1, N + 1, dtype=np.int64)   # :53

    def next(self) -> int:
        if self.sweeping == 1:
            self.idxr = self.rng.rand_range(self.N)
        elif self.sweeping == 2:
            self.idxr = self.idxr % self.N + 1
        else:
            if self.idx == self.N:
                self.ind = self.rng.randperm(self.N)
                self.idx = 1
            else:
                self.idx += 1
            self.idxr = int(self.ind[self.idx - 1])
        return self.idxr

    def take(self, k: int) -> np.ndarray:
        return np.array([self.next() for _ in range(k)], dtype=np.int64)


class LFinitoSweeper:
    """Finito_LFinito.jl:89-91: ``inds = randperm(d)`` before every sweep when
    sweeping == 3, otherwise the natural order (sweeping 1 is silently cyclic)."""

    def __init__(self, N: int, batch: int, sweeping: int, rng: HostRNG):
        self.N, self.r, self.sweeping, self.rng = N, batch, sweeping, rng
        self.d = n_batches(N, batch)
        self.inds = np.arange(1, self.d + 1, dtype=np.int64)

    def next(self) -> np.ndarray:
        if self.sweeping == 3:
            self.inds = self.rng.randperm(self.d)
        return self.inds


def csr(batches):
    """list of 1-based index arrays → (idx, ptr) for the ``*_steps`` ABI calls."""
    ptr = np.zeros(len(batches) + 1, dtype=np.int64)
    if batches:
        ptr[1:] = np.cumsum([len(b) for b in batches])
        idx = np.ascontiguousarray(np.concatenate(batches), dtype=np.int64)
    else:
        idx = np.zeros(0, dtype=np.int64)
    return idx, ptr


def interleaved_rows(N: int, block: int, world: int, rank: int) -> np.ndarray:
    """Global (0-based) row numbers of an interleaved shard, in local order: blocks rank, rank + world, … of `block` rows."""
    g = np.arange(N, dtype=np.int64)
    return g[(g // block) % world == rank]


def shard_rows(N: int, world: int, rank: int):
    """Contiguous row partition: rank k owns rows [k·N/G, (k+1)·N/G)."""
    return (rank * N) // world, ((rank + 1) * N) // world

"""Host-side mirror of the reference's solver API over libciao_cuda.

Same names, keyword arguments, defaults, iteration-count semantics and error
behaviour as the Julia front-ends (the Julia twin is julia/CIAOAlgorithmsCUDA.jl):

  SVRG(γ, maxit, verbose, freq, m, plus)                       SVRG/SVRG.jl:24-44
  SAGA(γ, maxit, verbose, freq, SAG_flag), SAG(...)            SAGA_SAG/SAGA.jl:24-42, 190-191
  Finito(γ, sweeping, LFinito, adaptive, minibatch, maxit, verbose, freq, α, tol, tol_b)   Finito/Finito.jl:32-64
  Proshi(γ, sweeping, minibatch, maxit, verbose, freq, α)      ProShI/ProShI.jl:18-40
  solver(x0; F, g, L, μ, N) -> (solution, num_iters)           SVRG.jl:46-84 etc.
  iterator(solver, x0; F, g, L, μ, N) -> iterable              SVRG.jl:132-147 etc.
  solution(state)                                              SVRG_basic.jl:99 etc.

Count semantics (SVRG.jl:70-79): the init state is iteration 1, so maxit = K
performs K−1 steps.  ``iterate`` returning ``nothing`` (missing stepsize data)
becomes a ``warnings.warn`` + an empty iteration.  Index draws stay on the host
(sampling.py).  F may be a list of operators.py objects, or a ``DeviceProblem``
whose rows are already resident in HBM.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib as L
from . import operators as ops
from .engine import Engine
from .sampling import AdaptiveSweeper, BatchSweeper, HostRNG, LFinitoSweeper, csr, n_batches

GLOBAL_RNG = HostRNG(0)   # the reference draws from Julia's global RNG


def seed(s):
    """Random.seed!(s) for the host index stream."""
    global GLOBAL_RNG
    GLOBAL_RNG = HostRNG(s)


class DeviceProblem:
    """F already resident on a GPU (e.g. Engine.gen_synthetic); replaces the list of f_i objects."""

    def __init__(self, engine: Engine):
        self.engine = engine


def _element_type(x0):
    """The reference is generic in the element type (test_lasso.jl:3 runs Float32/Float64/ComplexF32/ComplexF64).  The engine
    computes in fp64; real single-precision problems are accepted and their solutions handed back in the caller's type
    (`eltype(x) == T`, test_lasso.jl:74).  Complex-typed problems are accepted when their data is real (zero imaginary parts:
    exactly what test_lasso.jl builds for ComplexF32/ComplexF64, see _lib.f64arr) and get complex-typed solutions back."""
    dt = np.asarray(x0).dtype
    if np.issubdtype(dt, np.complexfloating):
        return np.dtype(np.complex64) if dt == np.complex64 else np.dtype(np.complex128)
    return np.dtype(np.float32) if dt == np.float32 else np.dtype(np.float64)


def _setup_engine(F, g, N, device=0, x0=None):
    out_dtype = _element_type(x0) if x0 is not None else np.dtype(np.float64)
    x0_cplx = x0 is not None and np.iscomplexobj(x0) and bool(np.any(np.asarray(x0).imag != 0))
    e = _setup_engine_fp64(F, g, N, device, None if x0 is None else int(np.size(x0)), x0_cplx)
    e.out_dtype = out_dtype
    if getattr(e, "complex_data", False) and not np.issubdtype(out_dtype, np.complexfloating):
        raise ops.UnsupportedOperator("complex F needs a complex x0 (the reference's state vectors take x0's element type)")
    return e


def _x0(e, x0):
    """x0 as the engine takes it: real fp64, or — genuinely complex problems — realified as (re, im) pairs."""
    return L.realify_vec(x0) if getattr(e, "complex_data", False) else x0


def _setup_engine_fp64(F, g, N, device=0, d=None, force_complex=False):
    if N is None:
        raise TypeError("keyword argument N is required (SVRG.jl:52 `N = N`)")
    if isinstance(F, DeviceProblem):
        e = F.engine
        if e.N != N:
            raise ValueError(f"DeviceProblem holds N = {e.N}, got N = {N}")
    else:
        if F is None:
            F = [ops.Zero()] * N          # F === nothing && (F = fill(ProximalOperators.Zero(), (N,)))   SVRG.jl:58
        packed = ops.pack_F(list(F), N, d, force_complex)
        e = Engine(device)
        if packed[0] == "rows":
            e.set_rows(packed[1], packed[2], packed[3], packed[4])
        elif packed[0] == "rowblocks":
            e.set_row_blocks(packed[1], packed[2], packed[3], packed[5], packed[4])
            e.complex_data = packed[6]
        else:
            e.set_blocks(packed[1], packed[2], packed[3], packed[4])
    e.set_reg(*ops.reg_params(g, getattr(e, "complex_data", False)))
    return e


class _State:
    """Device-resident state; host mirrors are persistent arrays refreshed on access, so
    ``solution(state) is state.<field>`` holds as in the reference (test_lasso.jl:154,185,221)."""
    _vecs = {}

    def __init__(self, engine):
        self.engine = engine
        self._dtype = getattr(engine, "out_dtype", np.dtype(np.float64))
        self._cplx = bool(getattr(engine, "complex_data", False))       # device vectors hold (re, im) pairs
        n_host = engine.d // 2 if self._cplx else engine.d
        self._host = {name: np.empty(n_host, dtype=self._dtype) for name in self._vecs}
        self._f64 = None if self._dtype == np.float64 else np.empty(engine.d)

    def __getattr__(self, name):
        vecs = type(self)._vecs
        if name in vecs:
            if self._cplx:
                self._host[name][:] = L.complexify_vec(self.engine.get_vec(vecs[name], self._f64))
                return self._host[name]
            if self._f64 is None:
                return self.engine.get_vec(vecs[name], self._host[name])
            self._host[name][:] = self.engine.get_vec(vecs[name], self._f64)   # fp64 on the device, the caller's type outside
            return self._host[name]
        raise AttributeError(name)


# ---------------------------------------------------------------------------------------------
# SVRG
class SVRG_basic_state(_State):
    _vecs = {"z_full": L.VEC_Z_FULL, "z": L.VEC_Z, "w": L.VEC_W, "av": L.VEC_AV}

    def __init__(self, engine, gamma, m):
        super().__init__(engine)
        self.γ = self.gamma = gamma
        self.m = m


class SVRG_basic_iterable:
    def __init__(self, F, g, x0, N, L_, mu, gamma, m, plus, rng=None, device=0):
        self.F, self.g, self.x0, self.N, self.L, self.μ, self.γ, self.m, self.plus = F, g, x0, N, L_, mu, gamma, m, plus
        self.rng, self.device = rng, device

    def _init(self):
        N = self.N
        m = N if self.m is None else self.m                                    # SVRG_basic.jl:33
        if self.γ is None:
            if self.plus:
                warnings.warn("provide a stepsize γ")                          # :36-38
                return None
            if self.L is None or self.μ is None:
                warnings.warn("smoothness or convexity parameter absent")      # :40-42
                return None
            L_M, mu_M = float(np.max(self.L)), float(np.max(self.μ))
            gamma = 1 / (10 * L_M)                                             # :46
            rho = (1 + 4 * L_M * gamma ** 2 * mu_M * (N + 1)) / (mu_M * gamma * N * (1 - 4 * L_M * gamma))
            if rho >= 1:
                warnings.warn("convergence condition violated...provide a stepsize!")
        else:
            gamma = self.γ
        e = _setup_engine(self.F, self.g, N, self.device, self.x0)
        e.svrg_init(_x0(e, self.x0), gamma, self.plus)                                 # :58-66
        return SVRG_basic_state(e, gamma, m)

    def __iter__(self):
        state = self._init()
        if state is None:
            return
        rng = self.rng or GLOBAL_RNG
        yield state
        while True:
            state.engine.svrg_epoch(rng.rand_vec(self.N, state.m))             # :73-92
            if self.plus:
                state.m *= 2                                                   # :93
            yield state


class SVRG:
    def __init__(self, gamma=None, maxit=10000, verbose=False, freq=1000, m=None, plus=False):
        assert gamma is None or gamma > 0
        assert maxit > 0 and freq > 0
        self.γ, self.maxit, self.verbose, self.freq, self.m, self.plus = gamma, maxit, verbose, freq, m, plus

    def _iterable(self, x0, F=None, g=None, L=None, mu=None, N=None, rng=None, device=0):
        return SVRG_basic_iterable(F, g, x0, N, L, mu, self.γ, self.m, self.plus, rng, device)

    def __call__(self, x0, **kw):
        maxit = self.maxit
        if self.plus and maxit > 25:                                           # SVRG.jl:62-65
            maxit = 25
            warnings.warn("exponential number of inner updates...reverted to 25 maximum iterations")
        return _drive(self, self._iterable(x0, **kw), maxit, lambda s: s.γ)


# ---------------------------------------------------------------------------------------------
# SAGA / SAG
class SAGA_basic_state(_State):
    _vecs = {"z": L.VEC_Z, "av": L.VEC_AV}

    def __init__(self, engine, gamma):
        super().__init__(engine)
        self.γ = self.gamma = gamma
        self.ind = 1

    @property
    def s(self):
        return self.engine.get_table_rows()


class SAGA_basic_iterable:
    def __init__(self, F, g, x0, N, L_, gamma, sag, rng=None, device=0):
        self.F, self.g, self.x0, self.N, self.L, self.γ, self.SAG = F, g, x0, N, L_, gamma, sag
        self.rng, self.device = rng, device

    def _init(self):
        if self.γ is None:
            if self.L is None:
                warnings.warn("smoothness parameter absent")                   # SAGA_basic.jl:30-32
                return None
            L_M = float(np.max(self.L))
            gamma = 1 / (16 * L_M) if self.SAG else 1 / (3 * L_M)              # :35
        else:
            gamma = self.γ
        e = _setup_engine(self.F, self.g, self.N, self.device, self.x0)
        e.saga_init(_x0(e, self.x0), gamma, self.SAG)                                  # :41-48
        return SAGA_basic_state(e, gamma)

    def steps(self, state, k):
        """k reference iterations in one persistent kernel; draws k × rand(1:N) first (:55)."""
        rng = self.rng or GLOBAL_RNG
        idx = rng.rand_vec(self.N, k)
        state.ind = int(idx[-1])
        state.engine.saga_steps(idx)

    def __iter__(self):
        state = self._init()
        if state is None:
            return
        yield state
        while True:
            self.steps(state, 1)
            yield state


class SAGA:
    def __init__(self, gamma=None, maxit=10000, verbose=False, freq=1000, SAG_flag=False):
        assert gamma is None or gamma > 0
        assert maxit > 0 and freq > 0
        self.γ, self.maxit, self.verbose, self.freq, self.SAG_flag = gamma, maxit, verbose, freq, SAG_flag

    def _iterable(self, x0, F=None, g=None, L=None, N=None, rng=None, device=0):
        return SAGA_basic_iterable(F, g, x0, N, L, self.γ, self.SAG_flag, rng, device)

    def __call__(self, x0, **kw):
        return _drive(self, self._iterable(x0, **kw), self.maxit, lambda s: s.γ)


def SAG(**kw):                                                                 # SAGA.jl:190-191
    return SAGA(SAG_flag=True, **kw)


# ---------------------------------------------------------------------------------------------
# Finito / MISO / DIAG (basic and low-memory)
def _finito_gammas(N, L_, gamma, alpha):
    """Finito_basic.jl:61-74 / Finito_LFinito.jl:51-63 / ProShI_basic.jl:61-74"""
    if gamma is None:
        if L_ is None:
            warnings.warn("--> smoothness parameter absent")
            return None
        if np.ndim(L_) == 0:
            return np.full(N, alpha * float(N) / float(L_))
        return alpha * float(N) / np.asarray(L_, dtype=np.float64)
    if np.ndim(gamma) == 0:
        return np.full(N, float(gamma))
    return np.asarray(gamma, dtype=np.float64)


class FINITO_basic_state(_State):
    _vecs = {"z": L.VEC_Z, "av": L.VEC_AV}

    def __init__(self, engine, gam, hat_gamma, sweeper):
        super().__init__(engine)
        self.γ = self.gamma = gam
        self.hat_γ = self.hat_gamma = hat_gamma
        self.sweeper = sweeper
        self.d = sweeper.d

    @property
    def s(self):
        return self.engine.get_table_rows()


class FINITO_basic_iterable:
    def __init__(self, F, g, x0, N, L_, gamma, sweeping, batch, alpha, rng=None, device=0):
        self.F, self.g, self.x0, self.N, self.L, self.γ = F, g, x0, N, L_, gamma
        self.sweeping, self.batch, self.α, self.rng, self.device = sweeping, batch, alpha, rng, device

    def _init(self):
        gam = _finito_gammas(self.N, self.L, self.γ, self.α)
        if gam is None:
            return None
        hat = 1 / np.sum(1 / gam)                                              # Finito_basic.jl:82
        e = _setup_engine(self.F, self.g, self.N, self.device, self.x0)
        e.finito_init(_x0(e, self.x0), gam, hat)                                       # :76-84
        return FINITO_basic_state(e, gam, hat, BatchSweeper(self.N, self.batch, self.sweeping, self.rng or GLOBAL_RNG))

    def steps(self, state, k):
        idx, bp = csr(state.sweeper.take(k))                                   # :96-108
        state.engine.finito_steps(idx, bp)                                     # :110-118

    def __iter__(self):
        state = self._init()
        if state is None:
            return
        yield state
        while True:
            self.steps(state, 1)
            yield state


class FINITO_LFinito_state(_State):
    _vecs = {"z": L.VEC_Z, "z_full": L.VEC_Z_FULL, "av": L.VEC_AV}

    def __init__(self, engine, gam, hat_gamma, sweeper):
        super().__init__(engine)
        self.γ = self.gamma = gam
        self.hat_γ = self.hat_gamma = hat_gamma
        self.sweeper = sweeper
        self.d = sweeper.d


class FINITO_LFinito_iterable(FINITO_basic_iterable):
    def _init(self):
        gam = _finito_gammas(self.N, self.L, self.γ, self.α)
        if gam is None:
            return None
        hat = 1 / np.sum(1 / gam)                                              # Finito_LFinito.jl:66
        e = _setup_engine(self.F, self.g, self.N, self.device, self.x0)
        e.lfinito_init(_x0(e, self.x0), gam, hat)                                      # :67-72
        return FINITO_LFinito_state(e, gam, hat, LFinitoSweeper(self.N, self.batch, self.sweeping, self.rng or GLOBAL_RNG))

    def steps(self, state, k):
        for _ in range(k):
            state.engine.lfinito_outer(state.sweeper.next(), self.batch)       # :78-103


class FINITO_adaptive_state(_State):
    """Finito_adaptive.jl:13-31.  γ and hat_γ change on the device during the linesearch: they are read back on access.
    ∇f is held as the scalars c_i (∇f_i(x_i) = c_i·a_i for the row models the engine supports)."""
    _vecs = {"z": L.VEC_Z, "av": L.VEC_AV}

    def __init__(self, engine, sweeper):
        super().__init__(engine)
        self.sweeper = sweeper

    @property
    def γ(self):
        return self.engine.finito_adaptive_get()[0]

    gamma = γ

    @property
    def hat_γ(self):
        return self.engine.finito_adaptive_get(gamma=False)[3]

    hat_gamma = hat_γ

    @property
    def fi_x(self):
        return self.engine.finito_adaptive_get(gamma=False, fi_x=True)[1]

    @property
    def backtracks(self):
        return self.engine.finito_adaptive_get(gamma=False)[4]

    @property
    def s(self):
        return self.engine.get_table_rows()


class FINITO_adaptive_iterable:
    """Finito_adaptive.jl:1-11; `tol` is carried but unused by the reference's iterate as well."""

    def __init__(self, F, g, x0, N, L_, tol, tol_b, sweeping, alpha, rng=None, device=0):
        self.F, self.g, self.x0, self.N, self.L, self.tol, self.tol_b = F, g, x0, N, L_, tol, tol_b
        self.sweeping, self.α, self.rng, self.device = sweeping, alpha, rng, device

    def _init(self):
        e = _setup_engine(self.F, self.g, self.N, self.device, self.x0)
        rng = self.rng or GLOBAL_RNG
        d = int(np.size(self.x0))
        # :59-99; the random restart of the stepsize estimate (:77-83) draws from the host's RNG, in the reference's order
        e.finito_adaptive_init(self.x0, self.α, self.tol_b, perturb=lambda i, t: rng.rand_pm(t, d))
        return FINITO_adaptive_state(e, AdaptiveSweeper(self.N, self.sweeping, self.rng or GLOBAL_RNG))

    def steps(self, state, k):
        """k steps in one persistent-kernel call; returns the number completed (the linesearch may end the iteration)."""
        done = state.engine.finito_adaptive_steps(state.sweeper.take(k))       # :101-160
        if done < k:
            warnings.warn("parameter `γ` became too small")                    # :125
        return done

    def __iter__(self):
        state = self._init()
        yield state
        while True:
            if self.steps(state, 1) < 1:
                return                                                         # `return nothing`
            yield state


class Finito:
    def __init__(self, gamma=None, sweeping=1, LFinito=False, adaptive=False, minibatch=(False, 1), maxit=10000,
                 verbose=False, freq=10000, alpha=0.999, tol=1e-8, tol_b=1e-9):
        assert gamma is None or np.min(gamma) > 0
        assert maxit > 0 and tol > 0 and tol_b > 0 and freq > 0
        self.γ, self.sweeping, self.LFinito, self.adaptive, self.minibatch = gamma, sweeping, LFinito, adaptive, minibatch
        self.maxit, self.verbose, self.freq, self.α, self.tol, self.tol_b = maxit, verbose, freq, alpha, tol, tol_b

    def _iterable(self, x0, F=None, g=None, L=None, N=None, rng=None, device=0):
        if self.LFinito:                                                       # Finito.jl:80-116
            cls = FINITO_LFinito_iterable
        elif self.adaptive:
            return FINITO_adaptive_iterable(F, g, x0, N, L, self.tol, self.tol_b, self.sweeping, self.α, rng, device)
        else:
            cls = FINITO_basic_iterable
        return cls(F, g, x0, N, L, self.γ, self.sweeping, self.minibatch[1], self.α, rng, device)

    def __call__(self, x0, **kw):
        return _drive(self, self._iterable(x0, **kw), self.maxit, lambda s: s.hat_γ)


# ---------------------------------------------------------------------------------------------
# ProShI
class Proshi_basic_state(_State):
    _vecs = {"z": L.VEC_Z, "av": L.VEC_AV}

    def __init__(self, engine, gam, hat_gamma, sweeper):
        super().__init__(engine)
        self.γ = self.gamma = gam
        self.hat_γ = self.hat_gamma = hat_gamma
        self.sweeper = sweeper
        self.d = sweeper.d
        self._s = np.empty((engine.N, engine.d))

    @property
    def s(self):
        return self.engine.get_table_rows(out=self._s)


class Proshi_basic_iterable(FINITO_basic_iterable):
    def _init(self):
        gam = _finito_gammas(self.N, self.L, self.γ, self.α)
        if gam is None:
            return None
        hat = float(np.sum(gam))                                               # ProShI_basic.jl:82
        e = _setup_engine(self.F, self.g, self.N, self.device, self.x0)
        e.proshi_init(self.x0, gam, hat)                                       # :76-86
        return Proshi_basic_state(e, gam, hat, BatchSweeper(self.N, self.batch, self.sweeping, self.rng or GLOBAL_RNG))

    def steps(self, state, k):
        idx, bp = csr(state.sweeper.take(k))
        state.engine.proshi_steps(idx, bp)                                     # :111-123


class Proshi:
    def __init__(self, gamma=None, sweeping=1, minibatch=(False, 1), maxit=10000, verbose=False, freq=10000, alpha=0.999):
        assert gamma is None or np.min(gamma) > 0
        assert maxit > 0 and freq > 0
        self.γ, self.sweeping, self.minibatch, self.maxit = gamma, sweeping, minibatch, maxit
        self.verbose, self.freq, self.α = verbose, freq, alpha

    def _iterable(self, x0, F=None, g=None, L=None, N=None, rng=None, device=0):
        return Proshi_basic_iterable(F, g, x0, N, L, self.γ, self.sweeping, self.minibatch[1], self.α, rng, device)

    def __call__(self, x0, **kw):
        return _drive(self, self._iterable(x0, **kw), self.maxit, lambda s: s.hat_γ)


# ---------------------------------------------------------------------------------------------
def solution(state):
    if isinstance(state, SVRG_basic_state):
        return state.z_full                                                    # SVRG_basic.jl:99
    if isinstance(state, Proshi_basic_state):
        state.engine.proshi_solution(state._s)                                 # ProShI_basic.jl:127-132 (mutates)
        return state._s
    return state.z                                                             # SAGA_basic.jl:71, Finito_basic.jl:123, Finito_LFinito.jl:105


def iterator(solver, x0, **kw):
    """SVRG.jl:132-147 etc. — maxit, verbose, freq of the solver are ignored here."""
    return solver._iterable(x0, **kw)


def _drive(solver, iterable, maxit, disp_field):
    """The driver loop `for (it, state) in enumerate(take(halt(iter, stop), maxit))` (SVRG.jl:70-79),
    with the K−1 steps between two prints fused into one persistent-kernel call where the iterable
    offers ``steps`` (same RNG calls in the same order)."""
    if hasattr(iterable, "steps"):
        state = iterable._init()
        if state is None:
            raise TypeError("solution(nothing): the iterator ended before its first state")  # MethodError upstream
        it = 1
        while it < maxit:
            nxt = min(maxit, (it // solver.freq + 1) * solver.freq) if solver.verbose else maxit
            done = iterable.steps(state, nxt - it)
            if done is not None and done < nxt - it:                            # the iterator returned `nothing` (adaptive Finito)
                it += done
                break
            it = nxt
            if solver.verbose and it % solver.freq == 0:
                print("%5d | %.3e  " % (it, disp_field(state)))
        if solver.verbose and it % solver.freq != 0:
            print("%5d | %.3e  " % (it, disp_field(state)))
        return solution(state), it
    num_iters, state_final = None, None
    for it, state in enumerate(iterable, start=1):
        if solver.verbose and it % solver.freq == 0:
            print("%5d | %.3e  " % (it, disp_field(state)))
        num_iters, state_final = it, state
        if it >= maxit:
            break
    if state_final is None:
        raise TypeError("solution(nothing): the iterator ended before its first state")
    if solver.verbose and num_iters % solver.freq != 0:
        print("%5d | %.3e  " % (num_iters, disp_field(state_final)))
    return solution(state_final), num_iters

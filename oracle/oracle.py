"""ctypes front-end of the CPU oracle (oracle/ciao_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs; never by the
product package.  Each method is a thin call into the C restatement, which
cites the reference file:line it follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libciao_oracle.so")

LOSS_LS, LOSS_LOGISTIC, LOSS_DIAGQUAD = 0, 1, 2
REG_ZERO, REG_NORML1, REG_INDBOX, REG_NORML1_PAIRS = 0, 1, 2, 3
SYN_LASSO, SYN_LOGISTIC, SYN_SHARING = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


class _Problem(C.Structure):
    _fields_ = [
        ("loss_kind", C.c_int32), ("reg_kind", C.c_int32),
        ("N", C.c_int64), ("d", C.c_int64), ("lda", C.c_int64),
        ("A", _dp), ("b", _dp), ("lam", _dp),
        ("box_lo", C.c_double), ("box_hi", C.c_double), ("eta", C.c_double),
        ("reg_lambda", C.c_double), ("reg_lo", _dp), ("reg_hi", _dp),
        ("reg_lo_s", C.c_double), ("reg_hi_s", C.c_double),
        ("M", C.c_int64),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ciao_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ciao_gen.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr) if os.path.exists(f))
    if force or stale:
        subprocess.check_call(["make", "-B", "-C", _HERE, "libciao_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_gradient.restype = C.c_double
        _lib.orc_finito_hat_gamma.restype = C.c_double
        _lib.orc_proshi_hat_gamma.restype = C.c_double
        _lib.orc_max_row_sqnorm.restype = C.c_double
        _lib.orc_reg_value.restype = C.c_double
    return _lib


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _i(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class Problem:
    """(1/N) Σ f_i(x) + g(x)   or   (1/N) Σ f_i(x_i) + g(Σ x_i)."""

    def __init__(self, loss_kind, A, b, lam=None, *, box=(-2.0, 2.0), eta=0.0, rows_per_component=1):
        """rows_per_component = M > 1: component i is the M×d block of rows i·M … i·M+M−1 (LeastSquares / Precompose(LogisticLoss)
        with an M×d matrix); A then has N·M rows, b N·M entries, lam N entries."""
        self.A = _f64(A)
        self.M = int(rows_per_component)
        self.N, self.d = self.A.shape[0] // self.M, self.A.shape[1]
        assert self.A.shape[0] == self.N * self.M
        self.b = _f64(b)
        self.lam = _f64(np.ones(self.N) if lam is None else np.broadcast_to(lam, (self.N,)))
        self.loss_kind = loss_kind
        self.p = _Problem()
        self.p.loss_kind = loss_kind
        self.p.N, self.p.d, self.p.lda = self.N, self.d, self.d
        self.p.A, self.p.b, self.p.lam = _d(self.A), _d(self.b), _d(self.lam)
        self.p.box_lo, self.p.box_hi, self.p.eta = box[0], box[1], eta
        self.p.M = self.M
        self.set_reg(REG_ZERO)

    def set_reg(self, kind, lam=0.0, lo=-np.inf, hi=np.inf):
        self.p.reg_kind = kind
        self.p.reg_lambda = lam
        self._lo = self._hi = None
        self.p.reg_lo = self.p.reg_hi = None
        if np.ndim(lo) > 0:
            self._lo = _f64(lo)
            self.p.reg_lo = _d(self._lo)
        else:
            self.p.reg_lo_s = float(lo)
        if np.ndim(hi) > 0:
            self._hi = _f64(hi)
            self.p.reg_hi = _d(self._hi)
        else:
            self.p.reg_hi_s = float(hi)
        return self

    @property
    def ref(self):
        return C.byref(self.p)

    # -- operator protocol --------------------------------------------------
    def gradient(self, i, x):
        y = np.empty(self.d)
        f = lib().orc_gradient(self.ref, C.c_int64(i), _d(_f64(x)), _d(y))
        return y, f

    def prox(self, x, gamma):
        y = np.empty(self.d)
        lib().orc_prox(self.ref, _d(y), _d(_f64(x)), C.c_double(gamma))
        return y

    def objective(self, x):
        f, g = C.c_double(), C.c_double()
        lib().orc_objective(self.ref, _d(_f64(x)), C.byref(f), C.byref(g))
        return f.value, g.value

    def full_gradient(self, x, scale=1.0):
        out = np.empty(self.d)
        lib().orc_full_gradient(self.ref, _d(_f64(x)), C.c_double(scale), _d(out))
        return out

    def full_gradient_omp(self, x, scale=1.0, nthreads=None):
        """All-cores variant (NOT the reference's algorithm, which is single-threaded): the labelled generous CPU baseline."""
        out = np.empty(self.d)
        nt = num_threads() if nthreads is None else int(nthreads)
        lib().orc_full_gradient_omp(self.ref, _d(_f64(x)), C.c_double(scale), _d(out), C.c_int(nt))
        return out

    def max_row_sqnorm(self):
        return lib().orc_max_row_sqnorm(self.ref)


class SVRGState:
    """SVRG_basic.jl:13-24 (state) + :30-96 (iterate)."""

    def __init__(self, prob: Problem, x0, gamma, m=None, plus=False):
        self.prob, self.gamma, self.plus = prob, float(gamma), bool(plus)
        d = prob.d
        self.m = prob.N if m is None else int(m)
        self.av, self.z, self.z_full, self.w = (np.empty(d) for _ in range(4))
        lib().orc_svrg_init(prob.ref, _d(_f64(x0)), _d(self.av), _d(self.z), _d(self.z_full), _d(self.w))

    def inner(self, idx1):
        idx1 = _i64(idx1)
        lib().orc_svrg_inner(self.prob.ref, C.c_double(self.gamma), _i(idx1), C.c_int64(len(idx1)),
                             _d(self.av), _d(self.z), _d(self.z_full), _d(self.w))

    def epoch(self, idx1):
        idx1 = _i64(idx1)
        assert len(idx1) == self.m
        lib().orc_svrg_epoch(self.prob.ref, C.c_double(self.gamma), C.c_int(self.plus), _i(idx1),
                             C.c_int64(len(idx1)), _d(self.av), _d(self.z), _d(self.z_full), _d(self.w))
        if self.plus:
            self.m *= 2

    def solution(self):
        return self.z_full


class SAGAState:
    """SAGA_basic.jl:11-20 + :26-68."""

    def __init__(self, prob: Problem, x0, gamma, sag=False):
        self.prob, self.gamma, self.sag = prob, float(gamma), bool(sag)
        self.s = np.empty((prob.N, prob.d))
        self.av, self.z = np.empty(prob.d), np.empty(prob.d)
        lib().orc_saga_init(prob.ref, _d(_f64(x0)), C.c_double(self.gamma), _d(self.s), _d(self.av), _d(self.z))

    def steps(self, idx1):
        idx1 = _i64(idx1)
        lib().orc_saga_steps(self.prob.ref, C.c_double(self.gamma), C.c_int(self.sag), _i(idx1),
                             C.c_int64(len(idx1)), _d(self.s), _d(self.av), _d(self.z))

    def solution(self):
        return self.z


def _csr(batches):
    ptr = np.zeros(len(batches) + 1, dtype=np.int64)
    for j, b in enumerate(batches):
        ptr[j + 1] = ptr[j] + len(b)
    idx = np.concatenate([np.asarray(b, dtype=np.int64) for b in batches]) if batches else np.zeros(0, np.int64)
    return _i64(idx), ptr


class FinitoState:
    """Finito_basic.jl:13-26 + :44-121 (index selection is the caller's)."""

    def __init__(self, prob: Problem, x0, gamma):
        self.prob = prob
        self.gamma = _f64(np.broadcast_to(gamma, (prob.N,)))
        self.hat_gamma = lib().orc_finito_hat_gamma(_d(self.gamma), C.c_int64(prob.N))
        self.s = np.empty((prob.N, prob.d))
        self.av, self.z = np.empty(prob.d), np.empty(prob.d)
        lib().orc_finito_init(prob.ref, _d(_f64(x0)), _d(self.gamma), C.c_double(self.hat_gamma),
                              _d(self.s), _d(self.av), _d(self.z))

    def steps(self, batches, ptr=None):
        """batches: list of 1-based index arrays, or (with ptr) an already flattened CSR pair."""
        idx, ptr = _csr(batches) if ptr is None else (_i64(batches), _i64(ptr))
        batches = range(len(ptr) - 1)
        lib().orc_finito_steps(self.prob.ref, _d(self.gamma), C.c_double(self.hat_gamma), _i(idx), _i(ptr),
                               C.c_int64(len(batches)), _d(self.s), _d(self.av), _d(self.z))

    def solution(self):
        return self.z


class FinitoAdaptiveState:
    """Finito_adaptive.jl:13-57 + :59-160 (index selection is the caller's).  Tables s (x_i) and gf (∇f_i(x_i))."""

    def __init__(self, prob: Problem, x0, alpha=0.999, tol_b=1e-9, perturb=None):
        """perturb(i, t) -> d-vector `rand(t * [-1, 1], size(x0))` (the caller's RNG), for the random restart of :77-83."""
        self.prob, self.alpha, self.tol_b = prob, float(alpha), float(tol_b)
        N, d = prob.N, prob.d
        self.s, self.gf = np.empty((N, d)), np.empty((N, d))
        self.fi_x, self.gamma = np.empty(N), np.empty(N)
        self.av, self.z = np.empty(d), np.empty(d)
        hg = C.c_double()
        x0 = _f64(x0)
        CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int64, _dp)

        def _cb(_user, i1, t, out):
            if perturb is None:
                return 1
            np.ctypeslib.as_array(out, shape=(d,))[:] = x0 + np.asarray(perturb(int(i1), int(t)), dtype=np.float64)
            return 0

        cb = CB(_cb)
        lib().orc_finito_adaptive_init_cb.restype = C.c_int
        rc = lib().orc_finito_adaptive_init_cb(prob.ref, _d(x0), C.c_double(self.alpha), _d(self.s), _d(self.gf),
                                               _d(self.fi_x), _d(self.gamma), C.byref(hg), _d(self.av), _d(self.z), cb, None)
        if rc != 0:
            raise ValueError("∇f_i(x0 + 1) == ∇f_i(x0) and no perturbation source for the random restart (Finito_adaptive.jl:77-83)")
        self.hat_gamma = hg.value
        self.backtracks = 0

    def steps(self, idx1):
        """Returns the number of steps completed (< len(idx1) ⇔ the reference's `return nothing`, :125-128)."""
        idx = _i64(idx1)
        hg, nbt = C.c_double(self.hat_gamma), C.c_int64()
        lib().orc_finito_adaptive_steps.restype = C.c_int64
        done = lib().orc_finito_adaptive_steps(self.prob.ref, C.c_double(self.alpha), C.c_double(self.tol_b), _i(idx),
                                               C.c_int64(len(idx)), _d(self.s), _d(self.gf), _d(self.fi_x), _d(self.gamma),
                                               C.byref(hg), _d(self.av), _d(self.z), C.byref(nbt))
        self.hat_gamma = hg.value
        self.backtracks += nbt.value
        return int(done)

    def solution(self):
        return self.z


class LFinitoState:
    """Finito_LFinito.jl:13-24 + :40-103."""

    def __init__(self, prob: Problem, x0, gamma, batch=1):
        self.prob, self.r = prob, int(batch)
        self.gamma = _f64(np.broadcast_to(gamma, (prob.N,)))
        self.hat_gamma = lib().orc_finito_hat_gamma(_d(self.gamma), C.c_int64(prob.N))
        self.nb = -(-prob.N // self.r)
        self.av, self.z, self.z_full = (np.empty(prob.d) for _ in range(3))
        lib().orc_lfinito_init(prob.ref, _d(_f64(x0)), C.c_double(self.hat_gamma), _d(self.av), _d(self.z),
                               _d(self.z_full))

    def outer(self, batch_order1):
        o = _i64(batch_order1)
        lib().orc_lfinito_outer(self.prob.ref, _d(self.gamma), C.c_double(self.hat_gamma), _i(o),
                                C.c_int64(len(o)), C.c_int64(self.r), _d(self.av), _d(self.z), _d(self.z_full))

    def solution(self):
        return self.z


class ProshiState:
    """ProShI_basic.jl:13-26 + :44-132."""

    def __init__(self, prob: Problem, x0, gamma):
        self.prob = prob
        self.gamma = _f64(np.broadcast_to(gamma, (prob.N,)))
        self.hat_gamma = lib().orc_proshi_hat_gamma(_d(self.gamma), C.c_int64(prob.N))
        self.s = np.empty((prob.N, prob.d))
        self.av, self.z = np.empty(prob.d), np.empty(prob.d)
        lib().orc_proshi_init(prob.ref, _d(_f64(x0)), _d(self.gamma), C.c_double(self.hat_gamma),
                              _d(self.s), _d(self.av), _d(self.z))

    def steps(self, batches, ptr=None):
        """batches: list of 1-based index arrays, or (with ptr) an already flattened CSR pair."""
        idx, ptr = _csr(batches) if ptr is None else (_i64(batches), _i64(ptr))
        batches = range(len(ptr) - 1)
        lib().orc_proshi_steps(self.prob.ref, _d(self.gamma), C.c_double(self.hat_gamma), _i(idx), _i(ptr),
                               C.c_int64(len(batches)), _d(self.s), _d(self.av), _d(self.z))

    def solution(self):
        """Mutates the table on every call, like ProShI_basic.jl:127-132."""
        lib().orc_proshi_solution(self.prob.ref, _d(self.gamma), _d(self.z), _d(self.s))
        return self.s


def num_threads():
    lib().orc_num_threads.restype = C.c_int
    return int(lib().orc_num_threads())


def full_gradient_synth(kind, N, d, seed, lam, x, scale=1.0, i0=0, n=None, nthreads=None):
    """scale·Σ ∇f_i(x) and Σ f_i(x) over rows [i0, i0+n) of the synthetic problem, rows regenerated on the fly on all host
    cores with long-double accumulation: the full-scale check of the CUDA pass (SURVEY.md §8d parity protocol)."""
    n = N - i0 if n is None else n
    out, fs = np.empty(d), C.c_double()
    nt = num_threads() if nthreads is None else int(nthreads)
    lib().orc_full_gradient_synth_omp(C.c_int(kind), C.c_int64(d), C.c_uint64(seed), C.c_double(lam), C.c_int64(i0), C.c_int64(n),
                                      _d(_f64(x)), C.c_double(scale), _d(out), C.byref(fs), C.c_int(nt))
    return out, fs.value


# -- synthetic inputs (bit-identical to the device generator) -----------------
def gen_rows(kind, d, seed, i0, n):
    A = np.empty((n, d))
    rhs = np.empty(n)
    lib().orc_gen_rows(C.c_int(kind), C.c_int64(d), C.c_uint64(seed), C.c_int64(i0), C.c_int64(n), _d(A), _d(rhs))
    return A, (None if kind == SYN_SHARING else rhs)


def gen_xtrue(kind, d, seed):
    x = np.empty(d)
    lib().orc_gen_xtrue(C.c_int(kind), C.c_int64(d), C.c_uint64(seed), _d(x))
    return x

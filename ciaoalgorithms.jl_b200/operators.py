"""The ProximalOperators.jl objects the reference's tests hand to the solvers,
as plain host-side descriptors.  They carry data only: all arithmetic happens
in libciao_cuda (there is no CPU implementation here).

  LeastSquares(A, b, lam)                 test/test_lasso.jl:53-54
  Precompose(LogisticLoss(y, mu), L, 1.0) test/test_logistic_l1.jl:36
  Sum(Quadratic(Q, q), SqrDistL2(IndBox(lo, hi), eta))   test/test_sharing.jl:18-22
  NormL1(lam), IndBox(lo, hi), Zero()     test_lasso.jl:59, test_sharing.jl:25, SVRG.jl:49

``pack_F`` is the shim's "F/g recognition" step: it turns
F = [f_1..f_N] into the dense arrays ``ciao_set_rows`` / ``ciao_set_blocks`` take
and rejects anything the engine does not cover *before* any device call.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib as L


@dataclass
class LeastSquares:
    A: np.ndarray
    b: np.ndarray
    lam: float = 1.0


@dataclass
class LogisticLoss:
    y: np.ndarray
    mu: float = 1.0


@dataclass
class Precompose:
    f: object
    L: np.ndarray
    mu: float = 1.0
    b: float = 0.0


@dataclass
class Quadratic:
    Q: np.ndarray
    q: np.ndarray


@dataclass
class IndBox:
    lo: object
    hi: object


@dataclass
class SqrDistL2:
    ind: IndBox
    lam: float = 1.0


class Sum:
    def __init__(self, *fs):
        self.fs = fs


@dataclass
class NormL1:
    lam: float = 1.0


class Zero:
    pass


class UnsupportedOperator(TypeError):
    pass


def pack_F(F, N, d=None, force_complex=False):
    """→ ("rows", loss_kind, A[N,d], b[N], scale[N])  |  ("rowblocks", loss_kind, A[N·M,d], b[N·M], scale[N], M, complex)
    |  ("blocks", Qdiag[N,n], qlin[N,n], (lo,hi), eta)."""
    if len(F) != N:
        raise ValueError(f"F has {len(F)} components, N = {N}")
    f0 = F[0]
    if all(isinstance(f, Zero) for f in F):
        # F === nothing → fill(Zero(), N) (SVRG.jl:58, SAGA.jl:55, Finito.jl:78): ∇f_i ≡ 0, f_i ≡ 0.  Least-squares rows with
        # a_i = 0, b_i = 0, λ_i = 0 give exactly that (every product is +0), so no extra loss kind is needed.
        if d is None:
            raise UnsupportedOperator("all-Zero F needs the dimension of x0")
        return ("rows", L.LOSS_LS, np.zeros((N, d)), np.zeros(N), np.zeros(N))
    if isinstance(f0, LeastSquares):
        # LeastSquares(A_i (m×d), b_i (m), λ_i).  m = 1 with real data is the reference's own case (test_lasso.jl:53) and takes the
        # tuned kernels; m > 1 (uniform) runs as M×d block components; genuinely complex data (non-zero imaginary parts in A, b or
        # x0) is realified — x as (re, im) pairs, every complex row as the real block [Re; Im] — into blocks of M = 2m rows.
        mats = []
        for f in F:
            if not isinstance(f, LeastSquares):
                raise UnsupportedOperator("F mixes operator kinds")
            Ai = np.asarray(f.A)
            mats.append(Ai.reshape(1, -1) if Ai.ndim < 2 else Ai)
        m, dc = mats[0].shape
        if any(Ai.shape != (m, dc) for Ai in mats) or any(np.size(f.b) != m for f in F):
            raise UnsupportedOperator("engine covers LeastSquares terms of one common shape m×d with m entries of b each")
        is_cplx = force_complex or any(np.iscomplexobj(Ai) and np.any(Ai.imag != 0) for Ai in mats) or \
            any(np.iscomplexobj(f.b) and np.any(np.asarray(f.b).imag != 0) for f in F)
        s = np.array([float(np.real(f.lam)) for f in F])
        if is_cplx:
            A = np.vstack([L.realify_rows(Ai) for Ai in mats])
            b = np.concatenate([L.realify_vec(np.asarray(f.b).reshape(-1)) for f in F])
            return ("rowblocks", L.LOSS_LS, A, b, s, 2 * m, True)
        A = np.vstack([L.f64arr(Ai) for Ai in mats])          # complex-typed real data is accepted (zero imaginary parts)
        b = np.concatenate([L.f64arr(np.asarray(f.b).reshape(-1)) for f in F])
        if m == 1:
            return ("rows", L.LOSS_LS, A, b, s)
        return ("rowblocks", L.LOSS_LS, A, b, s, m, False)
    if isinstance(f0, Precompose) and isinstance(f0.f, LogisticLoss):
        d = np.asarray(f0.L).reshape(1, -1).shape[1]
        A, y, mu = np.empty((N, d)), np.empty(N), np.empty(N)
        for i, f in enumerate(F):
            Li = np.asarray(f.L, dtype=np.float64)
            if not (isinstance(f, Precompose) and isinstance(f.f, LogisticLoss)) or Li.size != d \
                    or np.size(f.f.y) != 1 or np.any(np.asarray(f.b) != 0):
                raise UnsupportedOperator("engine covers Precompose(LogisticLoss([y_i], μ), 1×d row, 1.0) (test_logistic_l1.jl:36)")
            A[i], y[i], mu[i] = Li.reshape(-1), float(np.asarray(f.f.y).reshape(-1)[0]), float(f.f.mu)
        return ("rows", L.LOSS_LOGISTIC, A, y, mu)
    if isinstance(f0, Sum):
        quad = lambda f: next((g for g in f.fs if isinstance(g, Quadratic)), None)  # noqa: E731
        dist = lambda f: next((g for g in f.fs if isinstance(g, SqrDistL2)), None)  # noqa: E731
        q0, d0 = quad(f0), dist(f0)
        if q0 is None or d0 is None or len(f0.fs) != 2:
            raise UnsupportedOperator("engine covers Sum(Quadratic(diag), SqrDistL2(IndBox)) (test_sharing.jl:18-22)")
        n = np.asarray(q0.q).size
        Qd, ql = np.empty((N, n)), np.empty((N, n))
        box, eta = (float(d0.ind.lo), float(d0.ind.hi)), float(d0.lam)
        for i, f in enumerate(F):
            qi, di = quad(f), dist(f)
            Q = np.asarray(qi.Q, dtype=np.float64)
            if Q.ndim == 2:
                if np.any(Q - np.diag(np.diag(Q)) != 0):
                    raise UnsupportedOperator("engine covers diagonal Quadratic terms only")
                Q = np.diag(Q)
            if (float(di.ind.lo), float(di.ind.hi)) != box or float(di.lam) != eta:
                raise UnsupportedOperator("engine covers one shared SqrDistL2(IndBox(lo,hi), η) for all blocks")
            Qd[i], ql[i] = Q, np.asarray(qi.q, dtype=np.float64)
        return ("blocks", Qd, ql, box, eta)
    raise UnsupportedOperator(f"f_i of type {type(f0).__name__} is outside the engine's scope (no CPU fallback)")


def reg_params(g, complex_data=False):
    if g is None or isinstance(g, Zero):
        return (L.REG_ZERO,)
    if isinstance(g, NormL1):
        # complex data: sign(x)·max(0, |x| − γλ) on every complex entry = a group soft-threshold on the (re, im) pairs
        return (L.REG_NORML1_PAIRS if complex_data else L.REG_NORML1, float(g.lam))
    if complex_data:
        raise UnsupportedOperator(f"g of type {type(g).__name__} on complex data is outside the engine's scope (Zero, NormL1)")
    if isinstance(g, IndBox):
        return (L.REG_INDBOX, g.lo, g.hi)
    raise UnsupportedOperator(f"g of type {type(g).__name__} is outside the engine's scope (Zero, NormL1, IndBox)")

"""The reference's own test files, replayed through the host-side mirror of its API
(operators.py objects → solvers.py → C ABI → CUDA).  Each test cites the @testset it mirrors.
"""
import itertools

import numpy as np
import pytest

import fixtures
from ciaoalgorithms_jl_b200 import operators as ops
from ciaoalgorithms_jl_b200 import solvers as S
from ciaoalgorithms_jl_b200.sampling import HostRNG

pytestmark = pytest.mark.gpu
TOL = 1e-4


def lasso_problem():
    fx = fixtures.planted_lasso(0)
    N = fx["N"]
    F = [ops.LeastSquares(fx["A"][i:i + 1, :], fx["b"][i:i + 1], float(N)) for i in range(N)]   # test_lasso.jl:52-58
    return fx, F, ops.NormL1(fx["lam"])


def logistic_problem():
    fx = fixtures.logistic_l1()
    F = [ops.Precompose(ops.LogisticLoss(np.array([fx["y"][i]]), 1.0), fx["A"][i].reshape(1, -1), 1.0)
         for i in range(fx["N"])]                                                               # test_logistic_l1.jl:36
    return fx, F, ops.NormL1(fx["lam"])


def sharing_problem():
    fx = fixtures.sharing()
    box = ops.IndBox(-2.0, 2.0)
    F = [ops.Sum(ops.Quadratic(np.diag(fx["Qdiag"][i]), np.ones(fx["n"])), ops.SqrDistL2(box, fx["eta"]))
         for i in range(fx["N"])]                                                               # test_sharing.jl:18-23
    return fx, F, ops.IndBox(-np.inf, np.ones(fx["n"]))


# ---- test_lasso.jl -------------------------------------------------------------------------
@pytest.mark.parametrize("sweeping", [1, 2, 3])
def test_lasso_basic_finito(sweeping):                                          # :70-75
    fx, F, g = lasso_problem()
    x, it = S.Finito(maxit=1000, sweeping=sweeping)(fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"], rng=HostRNG(1))
    assert fx["cost"](x) - fx["f_star"] < TOL and it == 1000 and x.dtype == np.float64


@pytest.mark.parametrize("sweeping,batch,lfinito", [(2, 1, True), (3, 1, True), (1, 2, False), (2, 2, False), (3, 3, False),
                                                     (2, 2, True), (3, 3, True)])
def test_lasso_lfinito_and_minibatch(sweeping, batch, lfinito):                 # :78-125
    fx, F, g = lasso_problem()
    solver = S.Finito(maxit=1000, sweeping=sweeping, LFinito=lfinito, minibatch=(True, batch))
    x, _ = solver(fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"], rng=HostRNG(1))
    assert fx["cost"](x) - fx["f_star"] < TOL


@pytest.mark.parametrize("sweeping", [1, 2, 3])
def test_lasso_adaptive_finito(sweeping):                                       # :88-98
    fx, F, g = lasso_problem()
    solver = S.Finito(maxit=1000, tol=1e-5, sweeping=sweeping, adaptive=True)
    x, it = solver(fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"], rng=HostRNG(1))
    assert fx["cost"](x) - fx["f_star"] < TOL and it == 1000 and x.dtype == np.float64


def test_lasso_adaptive_iterator_and_early_end():                               # :143-157 with adaptive = true; Finito_adaptive.jl:124-127
    fx, F, g = lasso_problem()
    it = S.iterator(S.Finito(sweeping=2, adaptive=True), fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"], rng=HostRNG(1))
    assert it.x0 is fx["x0"]
    for state in itertools.islice(it, 3):
        assert S.solution(state) is state.z and len(state.γ) == fx["N"] and state.hat_γ > 0
    # γ_i < tol_b/N at the first step: the iterator yields its initial state and ends, the solver stops at iteration 1
    with pytest.warns(UserWarning, match="became too small"):
        x, n = S.Finito(maxit=50, adaptive=True, tol_b=1e300)(fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"], rng=HostRNG(1))
    assert n == 1


def test_lasso_float32_problem_keeps_its_element_type():                        # test_lasso.jl:3 (T = Float32), :74 eltype(x) == T
    fx, F, g = lasso_problem()
    N = fx["N"]
    F32 = [ops.LeastSquares(fx["A"][i:i + 1, :].astype(np.float32), fx["b"][i:i + 1].astype(np.float32), np.float32(N)) for i in range(N)]
    x0 = np.zeros(fx["n"], dtype=np.float32)
    for solver in (S.Finito(maxit=1000, sweeping=2), S.SAGA(maxit=1000, gamma=1 / (3 * fx["L"].max())), S.Finito(maxit=1000, adaptive=True)):
        x, _ = solver(x0, F=F32, g=g, L=fx["L"].astype(np.float32), N=N, rng=HostRNG(1))
        assert x.dtype == np.float32 and fx["cost"](x.astype(np.float64)) - fx["f_star"] < 1e-3   # data rounded to single precision
    x, _ = S.SAGA(gamma=1 / (3 * fx["L"].max()), maxit=50)(x0.astype(np.complex64) + 1j, F=F32, g=g, N=N, rng=HostRNG(1))
    assert x.dtype == np.complex64             # a complex x0 makes the problem complex: realified block path (test_gpu_blocks.py)


def test_lasso_scalar_gamma_and_scalar_L():                                     # :128-140
    fx, F, g = lasso_problem()
    N = fx["N"]
    x, _ = S.Finito(maxit=1000, gamma=N / fx["L"].max())(fx["x0"], F=F, g=g, L=fx["L"], N=N, rng=HostRNG(1))
    assert fx["cost"](x) - fx["f_star"] < TOL
    x, _ = S.Finito(maxit=1000)(fx["x0"], F=F, g=g, L=float(fx["L"].max()), N=N, rng=HostRNG(1))
    assert fx["cost"](x) - fx["f_star"] < TOL


@pytest.mark.parametrize("sweeping,lfinito", [(1, False), (2, False), (3, True)])
def test_lasso_finito_iterator(sweeping, lfinito):                              # :143-157
    fx, F, g = lasso_problem()
    solver = S.Finito(sweeping=sweeping, LFinito=lfinito)
    it = S.iterator(solver, fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"], rng=HostRNG(1))
    assert it.x0 is fx["x0"]
    for state in itertools.islice(it, 2):
        assert S.solution(state) is state.z and S.solution(state).dtype == np.float64


def test_lasso_svrg_and_svrg_plus_plus():                                       # :164-176
    fx, F, g = lasso_problem()
    gamma = 1 / (7 * fx["L"].max())
    x, it = S.SVRG(maxit=1000, gamma=gamma)(fx["x0"], F=F, g=g, N=fx["N"], rng=HostRNG(1))
    assert fx["cost"](x) - fx["f_star"] < TOL and it == 1000
    x, it = S.SVRG(maxit=16, gamma=gamma, m=1, plus=True)(fx["x0"], F=F, g=g, N=fx["N"], rng=HostRNG(1))
    assert fx["cost"](x) - fx["f_star"] < TOL and it == 16


def test_lasso_svrg_iterator_and_maxit_one():                                   # :179-193
    fx, F, g = lasso_problem()
    gamma = 1 / (7 * fx["L"].max())
    it = S.iterator(S.SVRG(gamma=gamma), fx["x0"], F=F, g=g, N=fx["N"], rng=HostRNG(1))
    assert it.x0 is fx["x0"]
    for state in itertools.islice(it, 2):
        assert S.solution(state) is state.z_full
    first = next(iter(it))
    x1, n = S.SVRG(gamma=gamma, maxit=1)(fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"])
    assert n == 1 and np.array_equal(S.solution(first), x1)                    # `==` upstream: bitwise


@pytest.mark.parametrize("sag", [False, True])
def test_lasso_saga_sag(sag):                                                   # :196-267
    fx, F, g = lasso_problem()
    make = S.SAG if sag else S.SAGA
    x, _ = make(maxit=10000 if sag else 1000)(fx["x0"], F=F, g=g, N=fx["N"], L=fx["L"], rng=HostRNG(1))
    assert fx["cost"](x) - fx["f_star"] < TOL
    gamma = 1 / ((16 if sag else 3) * fx["L"].max())
    x, _ = make(maxit=10000 if sag else 1000, gamma=gamma)(fx["x0"], F=F, g=g, N=fx["N"], rng=HostRNG(1))
    assert fx["cost"](x) - fx["f_star"] < TOL
    it = S.iterator(make(gamma=gamma), fx["x0"], F=F, g=g, N=fx["N"], rng=HostRNG(1))
    for state in itertools.islice(it, 2):
        assert S.solution(state) is state.z
    first = next(iter(it))
    x1, n = make(gamma=gamma, maxit=1)(fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"])
    assert n == 1 and np.array_equal(S.solution(first), x1)


def test_missing_stepsize_warns_and_ends():                                     # SVRG_basic.jl:36-42, SAGA_basic.jl:30-32
    fx, F, g = lasso_problem()
    with pytest.warns(UserWarning):
        assert list(S.iterator(S.SVRG(plus=True), fx["x0"], F=F, g=g, N=fx["N"])) == []
    with pytest.warns(UserWarning):
        assert list(S.iterator(S.SAGA(), fx["x0"], F=F, g=g, N=fx["N"])) == []
    with pytest.warns(UserWarning), pytest.raises(TypeError):
        S.Finito()(fx["x0"], F=F, g=g, N=fx["N"])                              # solution(nothing) upstream


# ---- test_logistic_l1.jl ---------------------------------------------------------------------
@pytest.mark.parametrize("sweeping", [1, 2, 3])
def test_logistic_finito_golden(sweeping):                                      # :54-59
    fx, F, g = logistic_problem()
    x, _ = S.Finito(maxit=9000, sweeping=sweeping)(fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"], rng=HostRNG(1))
    assert np.abs(x - fx["x_star"]).max() < TOL


def test_logistic_cyclic_solver_equals_iterator():                              # :111-122
    fx, F, g = logistic_problem()
    for lf in (False, True):
        solver = S.Finito(maxit=10, sweeping=2, LFinito=lf)
        x, _ = solver(fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"])
        last = None
        for last in itertools.islice(S.iterator(solver, fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"]), 10):
            pass
        assert np.array_equal(S.solution(last), x)


def test_logistic_svrg_golden():                                                # :126-138
    fx, F, g = logistic_problem()
    gamma = 1 / (10 * fx["L"].max())
    x, _ = S.SVRG(maxit=3000, gamma=gamma)(fx["x0"], F=F, g=g, N=fx["N"], rng=HostRNG(1))
    assert np.linalg.norm(x - fx["x_star"]) < TOL
    x, _ = S.SVRG(maxit=16, gamma=gamma, m=fx["N"], plus=True)(fx["x0"], F=F, g=g, N=fx["N"], rng=HostRNG(1))
    assert np.linalg.norm(x - fx["x_star"]) < TOL
    # γ = nothing with L and μ given: γ = 1/(10 L_max) and the Theorem 3.1 check (SVRG_basic.jl:44-52)
    with pytest.warns(UserWarning):
        x, _ = S.SVRG(maxit=50)(fx["x0"], F=F, g=g, L=fx["L"], mu=np.full(fx["N"], 1e-6), N=fx["N"], rng=HostRNG(1))
    assert x.shape == fx["x0"].shape


# ---- test_sharing.jl -------------------------------------------------------------------------
@pytest.mark.parametrize("sweeping,batch", [(1, 1), (2, 1), (3, 1), (1, 2), (2, 2), (3, 3)])
def test_sharing_proshi_golden(sweeping, batch):                                # :38-57
    fx, F, g = sharing_problem()
    solver = S.Proshi(maxit=1000, sweeping=sweeping, minibatch=(True, batch))
    x, _ = solver(fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"], rng=HostRNG(1))
    assert np.abs(x.sum(axis=0) - fx["sum_star"]).max() < TOL and x.shape == (fx["N"], fx["n"])


def test_sharing_iterator_solution_is_table():                                  # :76-84
    fx, F, g = sharing_problem()
    it = S.iterator(S.Proshi(sweeping=2), fx["x0"], F=F, g=g, L=fx["L"], N=fx["N"])
    for state in itertools.islice(it, 2):
        assert S.solution(state) is state.s


def test_unsupported_operator_is_rejected_before_any_device_call():
    fx, F, g = lasso_problem()
    with pytest.raises(ops.UnsupportedOperator):
        S.SAGA(gamma=0.1)(fx["x0"], F=[object()] * fx["N"], g=g, N=fx["N"])
    with pytest.raises(ops.UnsupportedOperator):
        S.Finito(adaptive=True)(fx["x0"], F=F, g=object(), L=fx["L"], N=fx["N"])


def test_default_F_is_all_zero_components():                                    # SVRG.jl:58, SAGA.jl:55: F === nothing → fill(Zero(), N)
    x0 = np.array([3.0, -0.2, 0.7, -5.0])
    soft = lambda v, t: np.sign(v) * np.maximum(np.abs(v) - t, 0.0)              # noqa: E731
    # SAGA with ∇f_i ≡ 0: z = prox((1−γ)x0, γ) (SAGA_basic.jl:48), then every step is z ← prox_g(z, γ)
    gam, lam = 0.25, 0.8
    x, it = S.SAGA(gamma=gam, maxit=4)(x0, F=None, g=ops.NormL1(lam), N=5, rng=HostRNG(1))
    want = soft((1 - gam) * x0, gam * lam)
    for _ in range(3):
        want = soft(want, gam * lam)
    assert it == 4 and np.allclose(x, want, rtol=0, atol=1e-15)
    # SVRG with g = Zero as well: nothing moves
    x, _ = S.SVRG(gamma=0.1, maxit=3, m=7)(x0, F=None, N=5, rng=HostRNG(1))
    assert np.allclose(x, x0, rtol=1e-15, atol=0)      # z_full = (Σ_m w)/m with w ≡ x0: equal up to the rounding of the mean


@pytest.mark.parametrize("T", [np.complex64, np.complex128])
def test_lasso_complex_typed_problem(T):                                           # test_lasso.jl:3 (T = ComplexF32, ComplexF64)
    """The reference's complex Lasso problems are complex only in their element type: C = rand(R, N, n) (:19), alpha, x_star and b
    carry zero imaginary parts, so its complex arithmetic never leaves the real axis.  The engine computes on the real parts and
    hands the solution back in the caller's type (:74 `eltype(x) == T`)."""
    fx, _, g = lasso_problem()
    N = fx["N"]
    R = np.float32 if T == np.complex64 else np.float64
    F = [ops.LeastSquares(fx["A"][i:i + 1, :].astype(T), fx["b"][i:i + 1].astype(T), R(N)) for i in range(N)]
    x0 = np.zeros(fx["n"], dtype=T)
    tol = 1e-3 if T == np.complex64 else TOL
    for solver in (S.Finito(maxit=1000, sweeping=2), S.SVRG(gamma=R(1 / (7 * fx["L"].max())), maxit=1000), S.SAGA(maxit=1000)):
        x, _ = solver(x0, F=F, g=g, L=fx["L"].astype(R), N=N, rng=HostRNG(1))
        assert x.dtype == T and np.all(x.imag == 0)
        assert fx["cost"](x.real.astype(np.float64)) - fx["f_star"] < tol
    # data that really leaves the real axis takes the realified block path (tests/test_gpu_blocks.py): complex solution, finite
    F[2] = ops.LeastSquares(fx["A"][2:3, :].astype(T) * (1 + 1j), fx["b"][2:3].astype(T), R(N))
    x, _ = S.SAGA(maxit=200, gamma=1 / (6 * fx["L"].max()))(x0, F=F, g=g, N=N, rng=HostRNG(1))
    assert x.dtype == T and np.all(np.isfinite(x)) and np.any(x.imag != 0)


def test_lasso_with_two_rows_per_component():                                   # test_lasso.jl:52-54 with A[2i-1:2i, :] instead of A[i:i, :]
    """LeastSquares terms that are 2×n blocks: (1/3)Σ_i (3/2)‖A_i x − b_i‖² is the same Lasso objective, so the planted optimum holds."""
    fx, _, g = lasso_problem()
    N, M = fx["N"] // 2, 2
    F = [ops.LeastSquares(fx["A"][M * i:M * (i + 1), :], fx["b"][M * i:M * (i + 1)], float(N)) for i in range(N)]
    Lc = np.array([np.linalg.norm(fx["A"][M * i:M * (i + 1), :], 2) ** 2 * N for i in range(N)])          # opnorm(tempA)^2 * N, :55
    for solver in (S.Finito(maxit=2000, sweeping=2), S.Finito(maxit=2000, sweeping=1, LFinito=True), S.SAGA(maxit=3000),
                   S.SVRG(gamma=1 / (7 * Lc.max()), maxit=500)):
        x, _ = solver(fx["x0"], F=F, g=g, L=Lc, N=N, rng=HostRNG(1))
        assert fx["cost"](x) - fx["f_star"] < TOL
    with pytest.raises(Exception):                                              # adaptive Finito: not for block components
        S.Finito(maxit=10, adaptive=True)(fx["x0"], F=F, g=g, L=Lc, N=N, rng=HostRNG(1))

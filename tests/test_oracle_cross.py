"""Per-step cross-check of the C oracle against a second, independent restatement of the reference's loops
(tests/second_restatement.py: numpy array statements written from the Julia sources).  The reference has no per-step
golden trajectories and cannot run here, so the per-step behaviour of the oracle is pinned by two
independent transcriptions agreeing to rounding on the reference's own test problems — after EVERY step, on every
state vector and table row, for every solver variant.  The only licensed difference is the summation order inside a
dot product / a sum over N (BLAS order is unspecified upstream), hence a tolerance of a few ulp-scale units rather
than bit equality.  CPU only.
"""
import numpy as np
import pytest

import fixtures
import second_restatement as R2
from oracle import oracle as orc
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, LFinitoSweeper

RTOL = 5e-14   # observed: ≤ 4e-15 on every state vector and table row after every step (20 of the 35 cases agree below 1e-15)


def close(a, b, what, scale=None):
    a, b = np.atleast_1d(np.asarray(a, float)), np.atleast_1d(np.asarray(b, float))
    fin = np.isfinite(b)
    assert np.array_equal(a[~fin], b[~fin]), f"{what}: non-finite entries differ"     # e.g. f_i = +Inf on both sides (exp overflow)
    if not fin.any():
        return
    a, b = a[fin], b[fin]
    scale = max(1.0, float(np.max(np.abs(b)))) if scale is None else scale
    err = float(np.max(np.abs(a - b))) / scale
    assert err <= RTOL, f"{what}: {err:.3e}"


def logistic():
    fx = fixtures.logistic_l1()
    p = orc.Problem(orc.LOSS_LOGISTIC, fx["A"], fx["y"], fx["mu"]).set_reg(orc.REG_NORML1, lam=fx["lam"])
    F = [R2.LogisticRow(fx["A"][i], fx["y"][i], fx["mu"][i]) for i in range(fx["N"])]
    return fx, p, F, R2.NormL1(fx["lam"])


def lasso(seed=0):
    fx = fixtures.planted_lasso(seed)
    p = orc.Problem(orc.LOSS_LS, fx["A"], fx["b"], fx["scale"]).set_reg(orc.REG_NORML1, lam=fx["lam"])
    F = [R2.LeastSquaresRow(fx["A"][i], fx["b"][i], fx["scale"][i]) for i in range(fx["N"])]
    return fx, p, F, R2.NormL1(fx["lam"])


def sharing():
    fx = fixtures.sharing()
    p = orc.Problem(orc.LOSS_DIAGQUAD, fx["Qdiag"], fx["qlin"], box=fx["box"], eta=fx["eta"])
    p.set_reg(orc.REG_INDBOX, lo=-np.inf, hi=fx["g_hi"])
    F = [R2.DiagQuadPlusSqrDist(fx["Qdiag"][i], fx["qlin"][i], fx["box"][0], fx["box"][1], fx["eta"]) for i in range(fx["N"])]
    return fx, p, F, R2.IndBox(-np.inf, fx["g_hi"])


PROBLEMS = {"logistic": logistic, "lasso": lasso}


# ----------------------------------------------------------------------------------------------------------
def test_operators_agree():
    for make in (logistic, lasso, sharing):
        fx, p, F, g = make()
        rs = np.random.RandomState(3)
        for _ in range(5):
            x = rs.randn(fx["n"]) * 2
            for i in range(fx["N"]):
                y, f = p.gradient(i, x)
                y2, f2 = F[i].gradient(x)
                close(y, y2, "gradient")
                close(f, f2, "value")
            for gam in (0.1, 1.0, 7.5):
                close(p.prox(x, gam), g.prox(x, gam), "prox")


@pytest.mark.parametrize("name", ["logistic", "lasso"])
@pytest.mark.parametrize("plus", [False, True])
def test_svrg_every_epoch(name, plus):
    fx, p, F, g = PROBLEMS[name]()
    N = fx["N"]
    gamma = 1 / (7 * np.max(fx["L"]))
    a = orc.SVRGState(p, fx["x0"], gamma, m=N, plus=plus)
    b = R2.SVRG(F, g, fx["x0"], gamma, plus=plus)
    close(a.av, b.av, "init av")
    rng, m = HostRNG(5), N
    for k in range(6):
        idx = rng.rand_vec(N, m)
        a.epoch(idx)
        b.epoch(idx)
        for nm in ("av", "z", "z_full", "w"):
            close(getattr(a, nm), getattr(b, nm), f"epoch {k} {nm}")
        if plus:
            m *= 2


@pytest.mark.parametrize("name", ["logistic", "lasso"])
@pytest.mark.parametrize("sag", [False, True])
def test_saga_every_step(name, sag):
    fx, p, F, g = PROBLEMS[name]()
    N = fx["N"]
    gamma = 1 / ((16 if sag else 3) * np.max(fx["L"]))
    a = orc.SAGAState(p, fx["x0"], gamma, sag=sag)
    b = R2.SAGA(F, g, fx["x0"], gamma, sag=sag)
    close(a.s, np.array(b.s), "init table")
    close(a.av, b.av, "init av")
    close(a.z, b.z, "init z")
    rng = HostRNG(7)
    for k in range(200):
        i = rng.rand_range(N)
        a.steps([i])
        b.step(i)
        close(a.z, b.z, f"step {k} z")
        close(a.av, b.av, f"step {k} av")
        close(a.s, np.array(b.s), f"step {k} table")


@pytest.mark.parametrize("name", ["logistic", "lasso"])
@pytest.mark.parametrize("sweeping,batch", [(1, 1), (2, 1), (3, 1), (1, 2), (2, 3), (3, 3)])
def test_finito_every_step(name, sweeping, batch):
    fx, p, F, g = PROBLEMS[name]()
    N = fx["N"]
    gam = 0.999 * N / fx["L"]
    a = orc.FinitoState(p, fx["x0"], gam)
    b = R2.Finito(F, g, fx["x0"], gam)
    close(a.hat_gamma, b.hat, "hat_gamma")
    close(a.s, np.array(b.s), "init table")
    close(a.av, b.av, "init av")
    close(a.z, b.z, "init z")
    sw = BatchSweeper(N, batch, sweeping, HostRNG(11))
    for k in range(150):
        rows = sw.next()
        a.steps([rows])
        b.step(list(rows))
        close(a.z, b.z, f"step {k} z")
        close(a.av, b.av, f"step {k} av")
        close(a.s, np.array(b.s), f"step {k} table")


@pytest.mark.parametrize("name", ["logistic", "lasso"])
@pytest.mark.parametrize("sweeping,batch", [(2, 1), (3, 1), (2, 3), (3, 2)])
def test_lfinito_every_outer_iteration(name, sweeping, batch):
    fx, p, F, g = PROBLEMS[name]()
    N = fx["N"]
    gam = 0.999 * N / fx["L"]
    a = orc.LFinitoState(p, fx["x0"], gam, batch)
    b = R2.LFinito(F, g, fx["x0"], gam, batch)
    close(a.av, b.av, "init av")
    sw = LFinitoSweeper(N, batch, sweeping, HostRNG(13))
    for k in range(40):
        order = sw.next()
        a.outer(order)
        b.outer(list(order))
        for nm in ("av", "z", "z_full"):
            close(getattr(a, nm), getattr(b, nm), f"outer {k} {nm}")


@pytest.mark.parametrize("sweeping,batch", [(1, 1), (2, 1), (3, 1), (1, 2), (2, 2), (3, 3)])
def test_proshi_every_step_and_solution(sweeping, batch):
    fx, p, F, g = sharing()
    N = fx["N"]
    gam = 0.999 * N / fx["L"]
    a = orc.ProshiState(p, fx["x0"], gam)
    b = R2.ProShI(F, g, fx["x0"], gam)
    close(a.hat_gamma, b.hat, "hat_gamma")
    close(a.s, np.array(b.s), "init table")
    close(a.av, b.av, "init av")
    close(a.z, b.z, "init z")
    sw = BatchSweeper(N, batch, sweeping, HostRNG(17))
    for k in range(300):
        rows = sw.next()
        a.steps([rows])
        b.step(list(rows))
        close(a.z, b.z, f"step {k} z")
        close(a.av, b.av, f"step {k} av")
        close(a.s, np.array(b.s), f"step {k} table")
    close(a.solution(), np.array(b.solution()), "solution")        # mutates both tables the same way
    close(a.solution(), np.array(b.solution()), "solution, second call")


@pytest.mark.parametrize("name", ["logistic", "lasso"])
@pytest.mark.parametrize("sweeping", [1, 2, 3])
def test_finito_adaptive_every_step_same_linesearch_decisions(name, sweeping):
    from ciaoalgorithms_jl_b200.sampling import AdaptiveSweeper
    fx, p, F, g = PROBLEMS[name]()
    N = fx["N"]
    a = orc.FinitoAdaptiveState(p, fx["x0"], 0.999, 1e-9)
    b = R2.FinitoAdaptive(F, g, fx["x0"], 0.999, 1e-9)
    close(a.gamma, b.gam, "init gamma")
    close(a.hat_gamma, b.hat, "init hat_gamma")
    close(a.av, b.av, "init av")
    close(a.z, b.z, "init z")
    sw = AdaptiveSweeper(N, sweeping, HostRNG(19))
    zmax = 0.0
    for k in range(300):
        i = sw.next()
        done = a.steps([i])
        ok = b.step(i)
        assert (done == 1) == ok, f"step {k}: one restatement stopped, the other did not"
        if not ok:
            break
        assert a.backtracks == b.backtracks, f"step {k}: {a.backtracks} vs {b.backtracks} stepsize reductions"
        close(a.gamma, b.gam, f"step {k} gamma")
        close(a.hat_gamma, b.hat, f"step {k} hat_gamma")
        close(a.z, b.z, f"step {k} z")
        close(a.av, b.av, f"step {k} av")
        close(a.s, np.array(b.s), f"step {k} table")
        # ∇f_i(x_i) and f_i(x_i) are functions of u = a_i·x_i: their rounding scale is that of u, and the iterates of the
        # logistic problem reach 2e4 before the linesearch has shrunk the stepsizes (relative agreement of z stays ≤ 1e-14)
        zmax = max(zmax, float(np.max(np.abs(b.z))))
        uscale = max(1.0, float(np.max(np.abs(fx["A"]))) ** 2 * zmax * fx["n"])
        close(a.gf, np.array(b.gf), f"step {k} gradient table", scale=uscale)
        close(a.fi_x, b.fi_x, f"step {k} f_i(x_i)", scale=uscale)
    assert b.backtracks > 0        # the linesearch was exercised


def test_finito_adaptive_random_restart_of_the_stepsize_estimate():
    """Finito_adaptive.jl:77-83: a component with ∇f_i(x0 + 1) == ∇f_i(x0) (here: rows whose entries sum to zero) gets a random
    ±t perturbation of x0, t doubling per retry, and γ_i from the t after the loop.  Both restatements are fed the same draws
    (sampling.JuliaRNG: `rand(t * [-1, 1], size(x0))` as the reference would draw them) in the same order."""
    from ciaoalgorithms_jl_b200.sampling import JuliaRNG
    N, d = 7, 6
    rs = np.random.default_rng(5)
    A = rs.standard_normal((N, d))
    A[1] = [1, -1, 2, -2, 3, -3]            # Σ_k a_k = 0: the +1 shift does not change a·x
    A[4] = [2, 2, -1, -1, -2, 0]
    b = rs.standard_normal(N)
    p = orc.Problem(orc.LOSS_LS, A, b, np.full(N, float(N))).set_reg(orc.REG_NORML1, lam=0.1)
    F = [R2.LeastSquaresRow(A[i], b[i], float(N)) for i in range(N)]
    x0 = np.zeros(d)                        # integer data: a_i·(x0 + 1) − a_i·x0 is exactly 0 in any summation order
    with pytest.raises(ValueError):
        orc.FinitoAdaptiveState(p, x0)
    ra, rb, calls = JuliaRNG(11), JuliaRNG(11), []

    def pert_a(i, t):
        calls.append((i, t))
        return ra.rand_pm(t, d)

    a = orc.FinitoAdaptiveState(p, x0, perturb=pert_a)
    b2 = R2.FinitoAdaptive(F, R2.NormL1(0.1), x0, perturb=lambda i, t: rb.rand_pm(t, d))
    assert [c[0] for c in calls if c[1] == 1] == [2, 5]          # the degenerate components, ascending, each starting at t = 1
    close(a.gamma, b2.gam, "gamma after the restart")
    close(a.hat_gamma, b2.hat, "hat_gamma")
    close(a.av, b2.av, "av")
    close(a.z, b2.z, "z")
    for k, i in enumerate([2, 5, 1, 2, 7, 5, 3]):
        assert a.steps([i]) == 1 and b2.step(i)
        close(a.z, b2.z, f"step {k} z")
        close(a.gamma, b2.gam, f"step {k} gamma")

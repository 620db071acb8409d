"""T2/T3 — the CUDA engine (through the C ABI) against the CPU oracle on the same
inputs and the same host-generated index sequences; golden vectors of the
reference's tests; bitwise run-to-run determinism; error behaviour.

Tolerances (BASELINE.json north_star): rel. objective ≤ 1e-8, ‖x − x_ref‖/‖x_ref‖ ≤ 1e-6
after K epochs.  The per-pass / few-step checks below use much tighter bounds
(the only differences are summation order and FMA contraction in the dots).
"""
import numpy as np
import pytest

import fixtures
from oracle import oracle as orc
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import CiaoError, Engine
from ciaoalgorithms_jl_b200.sampling import AdaptiveSweeper, BatchSweeper, HostRNG, LFinitoSweeper, csr

pytestmark = pytest.mark.gpu

REL_ITERATE = 1e-6     # north_star tolerance after K epochs
REL_OBJECTIVE = 1e-8
TIGHT = 1e-11          # single pass / short runs: reduction-order noise only


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


# ----------------------------------------------------------------------------
def make_rows(kind, N, d, seed, lam_reg, scale=None):
    """Synthetic row problem on both sides: oracle (host arrays) and engine (device generator)."""
    syn = orc.SYN_LASSO if kind == orc.LOSS_LS else orc.SYN_LOGISTIC
    A, rhs = orc.gen_rows(syn, d, seed, 0, N)
    sc = float(N) if (scale is None and kind == orc.LOSS_LS) else (1.0 if scale is None else scale)
    p = orc.Problem(kind, A, rhs, np.full(N, sc)).set_reg(orc.REG_NORML1, lam=lam_reg)
    e = Engine(0)
    e.gen_synthetic(L.SYNTH_LASSO if kind == orc.LOSS_LS else L.SYNTH_LOGISTIC, N, d, seed, scale=sc)
    e.set_reg(L.REG_NORML1, lam_reg)
    return p, e


def fixture_rows(which):
    if which == "lasso":
        fx = fixtures.planted_lasso(0)
        p = orc.Problem(orc.LOSS_LS, fx["A"], fx["b"], fx["scale"]).set_reg(orc.REG_NORML1, lam=fx["lam"])
        e = Engine(0)
        e.set_rows(L.LOSS_LS, fx["A"], fx["b"], fx["scale"])
    else:
        fx = fixtures.logistic_l1()
        p = orc.Problem(orc.LOSS_LOGISTIC, fx["A"], fx["y"], fx["mu"]).set_reg(orc.REG_NORML1, lam=fx["lam"])
        e = Engine(0)
        e.set_rows(L.LOSS_LOGISTIC, fx["A"], fx["y"], fx["mu"])
    e.set_reg(L.REG_NORML1, fx["lam"])
    return fx, p, e


# ----------------------------------------------------------------------------
# K1 / K9 / K10: streaming passes and the device generator
@pytest.mark.parametrize("kind", [orc.LOSS_LS, orc.LOSS_LOGISTIC])
@pytest.mark.parametrize("N,d", [(7, 3), (1000, 64), (999, 130), (4096, 1024), (1536, 4096), (300, 8192)])
def test_full_gradient_and_objective(kind, N, d):
    p, e = make_rows(kind, N, d, 0x5EED0000 + d, lam_reg=0.1, scale=1.0)
    x = np.random.default_rng(d).standard_normal(d) / np.sqrt(d)
    g_ref = p.full_gradient(x, 1.0 / N)
    g = e.full_gradient(x, 1.0 / N)
    assert rel(g, g_ref) < TIGHT
    f_ref, r_ref = p.objective(x)
    f, r = e.objective(x)
    assert abs(f - f_ref) <= 1e-12 * abs(f_ref) and abs(r - r_ref) <= 1e-13 * abs(r_ref)
    assert abs(e.max_row_sqnorm() - p.max_row_sqnorm()) <= 1e-12 * p.max_row_sqnorm()
    # bitwise reproducible (no floating-point atomics)
    assert np.array_equal(g, e.full_gradient(x, 1.0 / N))
    e.close()


def test_full_gradient_host_rows_match_generated_rows():
    """ciao_set_rows (host upload, lda > d) and the device generator produce the same problem."""
    N, d = 513, 96
    A, b = Engine.gen_host(L.SYNTH_LASSO, d, 77, 0, N)
    A2, b2 = orc.gen_rows(orc.SYN_LASSO, d, 77, 0, N)
    assert np.array_equal(A, A2) and np.array_equal(b, b2)
    wide = np.zeros((N, d + 5))
    wide[:, :d] = A
    x = np.linspace(-1, 1, d)
    with Engine(0) as e1, Engine(0) as e2:
        e1.set_rows(L.LOSS_LS, wide[:, :d], b, 2.5)           # strided view: lda = d + 5
        e2.gen_synthetic(L.SYNTH_LASSO, N, d, 77, scale=2.5)
        assert np.array_equal(e1.full_gradient(x), e2.full_gradient(x))


@pytest.mark.parametrize("threads,stages,ctas", [(128, 3, 2), (512, 2, 1), (256, 8, 1)])
def test_full_gradient_tuning_shapes(threads, stages, ctas):
    p, e = make_rows(orc.LOSS_LS, 2000, 1024, 5, lam_reg=0.1, scale=1.0)
    x = np.random.default_rng(1).standard_normal(1024)
    e.set_tuning(pass_threads=threads, pass_stages=stages, pass_ctas_per_sm=ctas)
    assert rel(e.full_gradient(x), p.full_gradient(x)) < TIGHT
    e.close()


# ----------------------------------------------------------------------------
# K3: SVRG / SVRG++
@pytest.mark.parametrize("kind,N,d,cluster", [(orc.LOSS_LS, 600, 64, 0), (orc.LOSS_LOGISTIC, 700, 256, 0),
                                               (orc.LOSS_LS, 512, 1024, 0), (orc.LOSS_LS, 384, 4096, 0),
                                               (orc.LOSS_LS, 384, 4096, 2), (orc.LOSS_LS, 300, 1024, 1), (orc.LOSS_LOGISTIC, 512, 1024, 2),
                                               (orc.LOSS_LS, 300, 4096, 16)])
@pytest.mark.parametrize("plus", [False, True])
def test_svrg_epochs(kind, N, d, cluster, plus):
    p, e = make_rows(kind, N, d, 0xABC + d, lam_reg=0.05 if kind == orc.LOSS_LS else 1.0 / N, scale=None)
    if cluster:
        e.set_tuning(seq_cluster=cluster, seq_threads=256 if cluster <= 2 else 0)
    Lmax = p.max_row_sqnorm() * (N if kind == orc.LOSS_LS else 0.25)
    gamma = 1 / (7 * Lmax)
    x0 = np.zeros(d) if kind == orc.LOSS_LS else np.ones(d)
    m0 = N // 4 if plus else N
    ref = orc.SVRGState(p, x0, gamma, m=m0, plus=plus)
    e.svrg_init(x0, gamma, plus)
    assert rel(e.get_vec(L.VEC_AV), ref.av) < TIGHT
    rng = HostRNG(3)
    m = m0
    for _ in range(3):
        idx = rng.rand_vec(N, m)
        ref.epoch(idx)
        e.svrg_epoch(idx)
        if plus:
            m *= 2
    assert rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-9
    assert rel(e.get_vec(L.VEC_W), ref.w) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    assert np.all(e.get_vec(L.VEC_Z) == 0.0)
    f_ref, f = sum(p.objective(ref.z_full)), sum(e.objective(e.get_vec(L.VEC_Z_FULL)))
    assert abs(f - f_ref) <= REL_OBJECTIVE * abs(f_ref)
    e.close()


@pytest.mark.parametrize("kind,N,d", [(orc.LOSS_LS, 384, 4096), (orc.LOSS_LOGISTIC, 500, 256)])
def test_svrg_lfinito_without_cached_coefficients(kind, N, d, monkeypatch):
    """CIAO_CACHE_CZ=0: the step recomputes a_i·z_full (two dot products per step) instead of reading c_i(z_full) from
    the dense cache the full-gradient pass leaves behind — the path multi-process and windowed passes take."""
    monkeypatch.setenv("CIAO_CACHE_CZ", "0")
    p, e = make_rows(kind, N, d, 0xC2 + d, lam_reg=0.05 if kind == orc.LOSS_LS else 1.0 / N)
    Lmax = p.max_row_sqnorm() * (N if kind == orc.LOSS_LS else 0.25)
    x0 = np.zeros(d) if kind == orc.LOSS_LS else np.ones(d)
    ref = orc.SVRGState(p, x0, 1 / (7 * Lmax), m=N, plus=False)
    e.svrg_init(x0, 1 / (7 * Lmax), False)
    rng = HostRNG(8)
    for _ in range(2):
        idx = rng.rand_vec(N, N)
        ref.epoch(idx)
        e.svrg_epoch(idx)
    assert rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    Li = np.sum(p.A * p.A, axis=1) * (N if kind == orc.LOSS_LS else 0.25)
    gam = 0.999 * N / Li
    refl = orc.LFinitoState(p, x0 + 0.1, gam, 4)
    e.lfinito_init(x0 + 0.1, gam, refl.hat_gamma)
    sw = LFinitoSweeper(N, 4, 3, HostRNG(2))
    for _ in range(2):
        order = sw.next()
        refl.outer(order)
        e.lfinito_outer(order, 4)
    assert rel(e.get_vec(L.VEC_Z), refl.z) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), refl.av) < 1e-8
    e.close()


def test_svrg_staged_indices_and_determinism():
    N, d = 800, 512
    p, e = make_rows(orc.LOSS_LS, N, d, 11, lam_reg=0.05)
    gamma = 1 / (7 * N * p.max_row_sqnorm())
    idx = HostRNG(5).rand_vec(N, N)
    outs = []
    for staged in (False, True, False):
        e.svrg_init(np.zeros(d), gamma, False)
        if staged:
            e.stage_indices(idx)
            e.svrg_epoch(None, N)
        else:
            e.svrg_epoch(idx)
        outs.append(e.get_vec(L.VEC_Z_FULL))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    e.close()


# ----------------------------------------------------------------------------
# K2 + K4: SAGA / SAG
@pytest.mark.parametrize("kind,N,d", [(orc.LOSS_LS, 300, 64), (orc.LOSS_LOGISTIC, 500, 1024), (orc.LOSS_LS, 256, 4096)])
@pytest.mark.parametrize("sag", [False, True])
@pytest.mark.parametrize("table_path", ["tma_ring", "ldg_fallback"])
def test_saga_steps(kind, N, d, sag, table_path, monkeypatch):
    # table rows staged in the shared-memory ring by TMA (default) or prefetched into registers (fallback when the
    # ring does not fit); the knob is read by ciao_create
    monkeypatch.setenv("CIAO_SEQ_TABLE_LDG", "1" if table_path == "ldg_fallback" else "0")
    p, e = make_rows(kind, N, d, 0x5A6A + d, lam_reg=0.05 if kind == orc.LOSS_LS else 1.0 / N)
    Lmax = p.max_row_sqnorm() * (N if kind == orc.LOSS_LS else 0.25)
    gamma = 1 / ((16 if sag else 3) * Lmax)
    x0 = np.full(d, 0.25)
    ref = orc.SAGAState(p, x0, gamma, sag=sag)
    e.saga_init(x0, gamma, sag)
    assert rel(e.get_table_rows(), ref.s) < TIGHT
    assert rel(e.get_vec(L.VEC_AV), ref.av) < TIGHT
    assert rel(e.get_vec(L.VEC_Z), ref.z) < TIGHT
    rng = HostRNG(9)
    idx = np.array([rng.rand_range(N) for _ in range(3 * N)], dtype=np.int64)
    idx[10:14] = idx[9]          # force back-to-back repeats: exercises the table hazard path
    idx[40] = idx[38]
    idx[68] = idx[60]            # repeats at the edge of the 8-step prefetch window of the TMA ring
    idx[89] = idx[80]
    idx[111] = idx[100]
    ref.steps(idx)
    e.saga_steps(idx[:N])
    e.saga_steps(idx[N:])
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    assert rel(e.get_table_rows(), ref.s) < 1e-9
    e.close()


# ----------------------------------------------------------------------------
# K2 + K5: Finito / MISO / DIAG
@pytest.mark.parametrize("kind,N,d", [(orc.LOSS_LS, 250, 64), (orc.LOSS_LOGISTIC, 401, 1024)])
@pytest.mark.parametrize("sweeping,batch", [(1, 1), (2, 1), (3, 1), (1, 7), (2, 16), (3, 5)])
@pytest.mark.parametrize("table_path", ["tma_ring", "ldg_fallback"])
def test_finito_steps(kind, N, d, sweeping, batch, table_path, monkeypatch):
    monkeypatch.setenv("CIAO_SEQ_TABLE_LDG", "1" if table_path == "ldg_fallback" else "0")
    p, e = make_rows(kind, N, d, 0xF1 + d, lam_reg=0.05 if kind == orc.LOSS_LS else 1.0 / N)
    A = p.A
    Li = np.sum(A * A, axis=1) * (N if kind == orc.LOSS_LS else 0.25)
    gam = 0.999 * N / Li
    x0 = np.full(d, 0.1)
    ref = orc.FinitoState(p, x0, gam)
    e.finito_init(x0, gam, ref.hat_gamma)
    assert rel(e.get_table_rows(), ref.s) < TIGHT
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-10
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-10
    batches = BatchSweeper(N, batch, sweeping, HostRNG(4)).take(2 * (-(-N // batch)) + 3)
    ref.steps(batches)
    idx, bp = csr(batches)
    e.finito_steps(idx, bp)
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    assert rel(e.get_table_rows(), ref.s) < 1e-9
    e.close()


# Finito adaptive (Finito_adaptive.jl): same linesearch decisions as the oracle, step for step
@pytest.mark.parametrize("kind,N,d", [(orc.LOSS_LS, 250, 64), (orc.LOSS_LS, 300, 1024), (orc.LOSS_LS, 200, 4096), (orc.LOSS_LOGISTIC, 401, 1024),
                                      (orc.LOSS_LS, 9, 5)])
@pytest.mark.parametrize("sweeping", [1, 2, 3])
def test_finito_adaptive_steps(kind, N, d, sweeping):
    p, e = make_rows(kind, N, d, 0xAD + d, lam_reg=0.05 if kind == orc.LOSS_LS else 1.0 / N)
    x0 = np.full(d, 0.1)
    ref = orc.FinitoAdaptiveState(p, x0, alpha=0.999, tol_b=1e-9)
    e.finito_adaptive_init(x0, 0.999, 1e-9)
    gam, fi_x, coef, hat, _ = e.finito_adaptive_get(True, True, True)
    assert rel(gam, ref.gamma) < 1e-12 and rel(fi_x, ref.fi_x) < 1e-12
    assert abs(hat - ref.hat_gamma) <= 1e-12 * ref.hat_gamma
    assert rel(e.get_table_rows(), ref.s) == 0.0
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-10 and rel(e.get_vec(L.VEC_Z), ref.z) < 1e-10
    idx = AdaptiveSweeper(N, sweeping, HostRNG(4)).take(max(2 * N + 7, 80))
    idx[20:23] = idx[19]            # immediate repeats and repeats inside the prefetch window: scalar history + table re-read
    idx[40] = idx[33]
    idx[60] = idx[49]
    K1 = N // 2 + 3
    assert ref.steps(idx) == len(idx)
    assert e.finito_adaptive_steps(idx[:K1]) == K1               # two calls: state (γ, γ̂, tables) carries over
    assert e.finito_adaptive_steps(idx[K1:]) == len(idx) - K1
    gam, fi_x, coef, hat, nbt = e.finito_adaptive_get(True, True, True)
    assert nbt == ref.backtracks                                 # identical linesearch decisions
    assert rel(gam, ref.gamma) < 1e-12                           # γ_i only ever changes by factors 0.8: same count per component
    assert abs(hat - ref.hat_gamma) <= 1e-12 * ref.hat_gamma
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    assert rel(e.get_table_rows(), ref.s) < 1e-9
    assert rel(fi_x, ref.fi_x) < 1e-9
    e.close()


def test_finito_adaptive_stops_when_gamma_too_small():
    N, d = 64, 128
    p, e = make_rows(orc.LOSS_LS, N, d, 0xAD0, lam_reg=0.05)
    x0 = np.full(d, 0.1)
    ref = orc.FinitoAdaptiveState(p, x0, alpha=0.999, tol_b=1e-9)
    thr_gamma = np.sort(ref.gamma)[N // 2]                       # about half of the components start below tol_b/N
    tol_b = float(thr_gamma) * N
    ref = orc.FinitoAdaptiveState(p, x0, alpha=0.999, tol_b=tol_b)
    e.finito_adaptive_init(x0, 0.999, tol_b)
    idx = AdaptiveSweeper(N, 1, HostRNG(7)).take(200)
    done_ref = ref.steps(idx)
    done = e.finito_adaptive_steps(idx)
    assert done == done_ref < 200                                # `return nothing` at the same step (Finito_adaptive.jl:124-127)
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    with pytest.raises(CiaoError):
        e.finito_adaptive_steps(np.array([0, 1], dtype=np.int64))  # out-of-range index is rejected before any update
    e.close()


# K6: LFinito
@pytest.mark.parametrize("kind,N,d", [(orc.LOSS_LS, 250, 64), (orc.LOSS_LOGISTIC, 401, 1024)])
@pytest.mark.parametrize("sweeping,batch", [(2, 1), (3, 1), (2, 16), (3, 7)])
def test_lfinito_outer(kind, N, d, sweeping, batch):
    p, e = make_rows(kind, N, d, 0x1F + d, lam_reg=0.05 if kind == orc.LOSS_LS else 1.0 / N)
    Li = np.sum(p.A * p.A, axis=1) * (N if kind == orc.LOSS_LS else 0.25)
    gam = 0.999 * N / Li
    x0 = np.full(d, 0.1)
    ref = orc.LFinitoState(p, x0, gam, batch)
    e.lfinito_init(x0, gam, ref.hat_gamma)
    assert rel(e.get_vec(L.VEC_AV), ref.av) < TIGHT
    sw = LFinitoSweeper(N, batch, sweeping, HostRNG(2))
    for _ in range(3):
        order = sw.next()
        ref.outer(order)
        e.lfinito_outer(order, batch)
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    assert rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    e.close()


# ----------------------------------------------------------------------------
# K7: ProShI
@pytest.mark.parametrize("N,n", [(3, 2), (200, 64), (333, 1024)])
@pytest.mark.parametrize("sweeping,batch", [(1, 1), (2, 1), (3, 1), (1, 5), (2, 8), (3, 3)])
def test_proshi_steps(N, n, sweeping, batch):
    if N == 3:
        fx = fixtures.sharing()
        Q, ql, box, eta, Li = fx["Qdiag"], fx["qlin"], fx["box"], fx["eta"], fx["L"]
        batch = min(batch, 3)
    else:
        Q, _ = orc.gen_rows(orc.SYN_SHARING, n, 0x5EED0005, 0, N)
        ql, box, eta = np.ones((N, n)), (-2.0, 2.0), 10.0 * N
        Li = np.abs(Q).max(axis=1) + eta
    p = orc.Problem(orc.LOSS_DIAGQUAD, Q, ql, box=box, eta=eta).set_reg(orc.REG_INDBOX, lo=-np.inf, hi=np.ones(n))
    gam = 0.999 * N / Li
    x0 = np.zeros(n)
    ref = orc.ProshiState(p, x0, gam)
    with Engine(0) as e:
        if N == 3:
            e.set_blocks(Q, ql, box, eta)
        else:
            e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005)
        e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
        e.proshi_init(x0, gam, ref.hat_gamma)
        assert rel(e.get_table_rows(), ref.s) < TIGHT
        assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-10
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
        batches = BatchSweeper(N, batch, sweeping, HostRNG(4)).take(3 * (-(-N // batch)) + 2)
        ref.steps(batches)
        idx, bp = csr(batches)
        e.proshi_steps(idx, bp)
        assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-9
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-8
        assert rel(e.get_table_rows(), ref.s) < 1e-9
        # solution() mutates the table on every call (ProShI_basic.jl:127-132)
        s1 = ref.solution().copy()
        out = np.empty((N, n))
        e.proshi_solution(out)
        assert rel(out, s1) < 1e-9
        assert rel(e.table_colsum(), s1.sum(axis=0)) < 1e-9
        e.proshi_solution(out)
        assert rel(out, ref.solution()) < 1e-9


@pytest.mark.parametrize("N,n,batch", [(700, 100, 64), (2000, 1024, 256), (1500, 36, 700)])
@pytest.mark.parametrize("sweeping", [1, 2, 3])
def test_proshi_parallel_minibatch_kernel(N, n, batch, sweeping):
    """Batches of ≥ 64 blocks take the column-sliced parallel kernel (proshi_batch_kernel)."""
    Q, _ = orc.gen_rows(orc.SYN_SHARING, n, 0x5EED0005, 0, N)
    ql, box, eta = np.ones((N, n)), (-2.0, 2.0), 10.0 * N
    hi = np.linspace(0.5, 1.5, n)                                        # vector bounds for g = IndBox(-Inf, hi)
    p = orc.Problem(orc.LOSS_DIAGQUAD, Q, ql, box=box, eta=eta).set_reg(orc.REG_INDBOX, lo=-np.inf, hi=hi)
    gam = 0.999 * N / (np.abs(Q).max(axis=1) + eta)
    ref = orc.ProshiState(p, np.zeros(n), gam)
    with Engine(0) as e:
        e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005)
        e.set_reg(L.REG_INDBOX, -np.inf, hi)
        e.proshi_init(np.zeros(n), gam, ref.hat_gamma)
        batches = BatchSweeper(N, batch, sweeping, HostRNG(8)).take(3 * (-(-N // batch)) + 1)
        ref.steps(batches)
        idx, bp = csr(batches)
        e.proshi_steps(idx, bp)
        assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-9
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-8
        assert rel(e.get_table_rows(), ref.s) < 1e-9
        z1 = e.get_vec(L.VEC_Z)
        e.proshi_init(np.zeros(n), gam, ref.hat_gamma)                   # bitwise reproducible
        e.proshi_steps(idx, bp)
        assert np.array_equal(z1, e.get_vec(L.VEC_Z))


# ----------------------------------------------------------------------------
# golden vectors of the reference's tests, reached by the CUDA engine
def test_golden_logistic_finito_svrg_saga():
    fx, p, e = fixture_rows("logistic")
    N, x0 = fx["N"], fx["x0"]
    gam = 0.999 * N / fx["L"]
    hat = 1 / np.sum(1 / gam)
    for sweeping in (1, 2, 3):                                       # test_logistic_l1.jl:54-59
        e.finito_init(x0, gam, hat)
        idx, bp = csr(BatchSweeper(N, 1, sweeping, HostRNG(1)).take(8999))
        e.finito_steps(idx, bp)
        assert np.abs(e.get_vec(L.VEC_Z) - fx["x_star"]).max() < 1e-4
    e.lfinito_init(x0, gam, hat)                                     # :62-68
    for _ in range(2000):
        e.lfinito_outer(np.arange(1, N + 1), 1)
    assert np.abs(e.get_vec(L.VEC_Z) - fx["x_star"]).max() < 1e-4
    gamma = 1 / (10 * fx["L"].max())                                 # :126-131
    e.svrg_init(x0, gamma, False)
    rng = HostRNG(1)
    for _ in range(3000):
        e.svrg_epoch(rng.rand_vec(N, N))
    assert np.linalg.norm(e.get_vec(L.VEC_Z_FULL) - fx["x_star"]) < 1e-4
    e.saga_init(x0, 1 / (3 * fx["L"].max()), False)                  # :160-164
    e.saga_steps(HostRNG(1).rand_vec(N, 8999))
    assert np.linalg.norm(e.get_vec(L.VEC_Z) - fx["x_star"]) < 5e-3
    e.close()


def test_golden_lasso_all_solvers():
    fx, p, e = fixture_rows("lasso")
    N, x0, cost, fs = fx["N"], fx["x0"], fx["cost"], fx["f_star"]
    gam = 0.999 * N / fx["L"]
    hat = 1 / np.sum(1 / gam)
    for sweeping, batch in [(1, 1), (2, 1), (3, 1), (1, 2), (2, 2), (3, 3)]:   # test_lasso.jl:70-75, 101-111
        e.finito_init(x0, gam, hat)
        idx, bp = csr(BatchSweeper(N, batch, sweeping, HostRNG(1)).take(999))
        e.finito_steps(idx, bp)
        assert cost(e.get_vec(L.VEC_Z)) - fs < 1e-4
    for sweeping, batch in [(2, 1), (3, 1), (2, 2), (3, 3)]:                    # :78-85, 114-125
        e.lfinito_init(x0, gam, hat)
        sw = LFinitoSweeper(N, batch, sweeping, HostRNG(1))
        for _ in range(999):
            e.lfinito_outer(sw.next(), batch)
        assert cost(e.get_vec(L.VEC_Z)) - fs < 1e-4
    gamma = 1 / (7 * fx["L"].max())                                             # :164-176
    e.svrg_init(x0, gamma, False)
    rng = HostRNG(1)
    for _ in range(999):
        e.svrg_epoch(rng.rand_vec(N, N))
    assert cost(e.get_vec(L.VEC_Z_FULL)) - fs < 1e-4
    e.svrg_init(x0, gamma, True)
    m = 1
    for _ in range(15):
        e.svrg_epoch(rng.rand_vec(N, m))
        m *= 2
    assert cost(e.get_vec(L.VEC_Z_FULL)) - fs < 1e-4
    e.saga_init(x0, 1 / (3 * fx["L"].max()), False)                             # :199-203
    e.saga_steps(rng.rand_vec(N, 999))
    assert cost(e.get_vec(L.VEC_Z)) - fs < 1e-4
    e.saga_init(x0, 1 / (16 * fx["L"].max()), True)                             # :236-240
    e.saga_steps(rng.rand_vec(N, 9999))
    assert cost(e.get_vec(L.VEC_Z)) - fs < 1e-4
    e.close()


def test_golden_sharing_proshi():
    fx = fixtures.sharing()
    N = fx["N"]
    gam = 0.999 * N / fx["L"]
    with Engine(0) as e:
        e.set_blocks(fx["Qdiag"], fx["qlin"], fx["box"], fx["eta"])
        e.set_reg(L.REG_INDBOX, -np.inf, fx["g_hi"])
        for sweeping, batch in [(1, 1), (2, 1), (3, 1), (1, 2), (2, 2), (3, 3)]:  # test_sharing.jl:38-57
            e.proshi_init(fx["x0"], gam, float(np.sum(gam)))
            idx, bp = csr(BatchSweeper(N, batch, sweeping, HostRNG(1)).take(999))
            e.proshi_steps(idx, bp)
            e.proshi_solution(None)
            assert np.abs(e.table_colsum() - fx["sum_star"]).max() < 1e-4


# ----------------------------------------------------------------------------
# end-to-end tolerance of the north star after K epochs at a reduced scale
def test_svrg_pp_k_epochs_tolerance():
    N, d = 4096, 4096
    p, e = make_rows(orc.LOSS_LS, N, d, 0x5EED0003, lam_reg=N / 100.0)
    gamma = 1 / (7 * N * p.max_row_sqnorm())
    ref = orc.SVRGState(p, np.zeros(d), gamma, m=N // 16, plus=True)
    e.svrg_init(np.zeros(d), gamma, True)
    rng, m = HostRNG(0x1D0003), N // 16
    for _ in range(5):
        idx = rng.rand_vec(N, m)
        ref.epoch(idx)
        e.svrg_epoch(idx)
        m *= 2
    x, x_ref = e.get_vec(L.VEC_Z_FULL), ref.z_full
    assert rel(x, x_ref) < REL_ITERATE
    f, f_ref = sum(e.objective(x)), sum(p.objective(x_ref))
    assert abs(f - f_ref) <= REL_OBJECTIVE * abs(f_ref)
    assert f_ref < sum(p.objective(np.zeros(d)))      # it actually descended
    e.close()


# ----------------------------------------------------------------------------
# error behaviour at the boundary (reference: @warn + `return nothing`)
def test_errors():
    with Engine(0) as e:
        with pytest.raises(CiaoError) as ei:
            e.svrg_init(np.zeros(4), 0.1)
        assert ei.value.code == -3                                   # no problem set
        e.gen_synthetic(L.SYNTH_LASSO, 64, 16, 1)
        with pytest.raises(CiaoError):
            e.svrg_epoch(np.array([1, 2, 3], dtype=np.int64))        # epoch before init
        with pytest.raises(CiaoError):
            e.svrg_init(np.zeros(16), -1.0)                          # γ ≤ 0 (SVRG.jl:39)
        e.svrg_init(np.zeros(16), 1e-3)
        e.svrg_epoch(np.array([1, 65, 3], dtype=np.int64))           # 65 > N: flagged, memory-safe
        with pytest.raises(CiaoError) as ei:
            e.sync()
        assert ei.value.code == -1
        with pytest.raises(CiaoError):
            e.proshi_steps(np.array([1], dtype=np.int64), np.array([0, 1], dtype=np.int64))  # wrong problem kind
        with pytest.raises(CiaoError):
            e.set_reg(7)


# ----------------------------------------------------------------------------
# edge cases: empty / ragged inputs, tiny shapes, heavy index repetition
def test_edge_cases_empty_and_tiny():
    # N = 1, d = 1
    p = orc.Problem(orc.LOSS_LS, np.array([[2.0]]), np.array([1.0]), np.array([1.0])).set_reg(orc.REG_NORML1, lam=0.01)
    with Engine(0) as e:
        e.set_rows(L.LOSS_LS, np.array([[2.0]]), np.array([1.0]), 1.0)
        e.set_reg(L.REG_NORML1, 0.01)
        ref = orc.SAGAState(p, np.zeros(1), 0.05)
        e.saga_init(np.zeros(1), 0.05, False)
        e.saga_steps(np.zeros(0, dtype=np.int64))                    # K = 0: nothing happens
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-15
        idx = np.ones(50, dtype=np.int64)                            # the same row 50 times: every step is a table hazard
        ref.steps(idx)
        e.saga_steps(idx)
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-13 and rel(e.get_table_rows(), ref.s) < 1e-13
        gam = np.array([0.3])
        reff = orc.FinitoState(p, np.zeros(1), gam)
        e.finito_init(np.zeros(1), gam, reff.hat_gamma)
        e.finito_steps(np.zeros(0, dtype=np.int64), np.zeros(1, dtype=np.int64))          # zero batches
        batches = [np.array([1]), np.zeros(0, dtype=np.int64), np.array([1])]              # an empty batch in the middle
        reff.steps(batches)
        e.finito_steps(*csr(batches))
        assert rel(e.get_vec(L.VEC_Z), reff.z) < 1e-13
        e.svrg_init(np.zeros(1), 0.05, False)
        with pytest.raises(CiaoError):
            e.svrg_epoch(np.zeros(0, dtype=np.int64), 0)             # m must be positive


@pytest.mark.parametrize("N,d", [(37, 5), (64, 2), (129, 1027)])
def test_ragged_shapes_all_algorithms(N, d):
    """d not a multiple of 4 (zero-padded columns), remainder batches, N not a multiple of anything."""
    p, e = make_rows(orc.LOSS_LOGISTIC, N, d, 0xE0 + d, lam_reg=1.0 / N)
    Li = 0.25 * np.sum(p.A * p.A, axis=1)
    gam = 0.999 * N / Li
    x0 = np.linspace(-0.5, 0.5, d)
    rng = HostRNG(6)
    ref = orc.SVRGState(p, x0, 1 / (10 * Li.max()))
    e.svrg_init(x0, 1 / (10 * Li.max()), False)
    for _ in range(2):
        idx = rng.rand_vec(N, N)
        ref.epoch(idx)
        e.svrg_epoch(idx)
    assert rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-10
    reff = orc.FinitoState(p, x0, gam)
    e.finito_init(x0, gam, reff.hat_gamma)
    batches = BatchSweeper(N, 5, 3, rng).take(3 * (-(-N // 5)))
    reff.steps(batches)
    e.finito_steps(*csr(batches))
    assert rel(e.get_vec(L.VEC_Z), reff.z) < 1e-10 and rel(e.get_table_rows(), reff.s) < 1e-10
    refl = orc.LFinitoState(p, x0, gam, 7)
    e.lfinito_init(x0, gam, refl.hat_gamma)
    sw = LFinitoSweeper(N, 7, 3, rng)
    for _ in range(2):
        o = sw.next()
        refl.outer(o)
        e.lfinito_outer(o, 7)
    assert rel(e.get_vec(L.VEC_Z), refl.z) < 1e-10
    e.close()


# ----------------------------------------------------------------------------
# BASELINE.json's full size (C3: N = 2^22, d = 4096, 137.6 GB of row records) through size-independent properties
def test_full_scale_properties():
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    if free < 150e9:
        pytest.skip("needs a 180 GB B200")
    N, d = 1 << 22, 4096
    with Engine(0) as e:
        e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N))
        e.set_reg(L.REG_NORML1, N / 100.0)
        rs = np.random.default_rng(0)
        x, y = rs.standard_normal(d) * 1e-3, rs.standard_normal(d) * 1e-3
        gx, gy = e.full_gradient(x, 1.0 / N), e.full_gradient(y, 1.0 / N)
        # (1) bitwise run-to-run determinism of the 296-CTA streaming pass
        assert np.array_equal(gx, e.full_gradient(x, 1.0 / N))
        # (2) the least-squares gradient is affine in x
        gm = e.full_gradient(0.25 * x + 0.75 * y, 1.0 / N)
        assert rel(gm, 0.25 * gx + 0.75 * gy) < 1e-10
        # (3) a sum of row-window passes is the full pass (what the sharded multi-GPU pass relies on)
        acc = np.zeros(d)
        for k in range(4):
            e.set_pass_window(k * N // 4, N // 4)
            acc += e.full_gradient(x, 1.0 / N)
        e.set_pass_window(0, 0)
        assert rel(acc, gx) < 1e-12
        # (4) the planted model: the gradient at x_true is the noise term only, orders of magnitude below the one at 0
        xt = orc.gen_xtrue(orc.SYN_LASSO, d, 0x5EED0003)
        assert np.linalg.norm(e.full_gradient(xt, 1.0 / N)) < 1e-3 * np.linalg.norm(e.full_gradient(np.zeros(d), 1.0 / N))
        # (5) rows regenerated on the host agree with what the device holds: spot-check the objective on a row window
        A, b = Engine.gen_host(L.SYNTH_LASSO, d, 0x5EED0003, N - 1000, 1000)
        e.set_pass_window(N - 1000, 1000)
        f_dev = e.objective(x)[0] * N
        e.set_pass_window(0, 0)
        f_host = float(np.sum(0.5 * N * (A @ x - b) ** 2))
        assert abs(f_dev - f_host) <= 1e-11 * f_host
        # (6) SVRG++ outer iterations descend monotonically from x0 = 0 and are reproducible bit for bit
        gamma = 1.0 / (7.0 * N * e.max_row_sqnorm())
        idx = HostRNG(1).rand_vec(N, N // 16)
        outs = []
        for _ in range(2):
            e.svrg_init(np.zeros(d), gamma, True)
            f0 = sum(e.objective(np.zeros(d)))
            e.svrg_epoch(idx)
            z1 = e.get_vec(L.VEC_Z_FULL)
            f1 = sum(e.objective(z1))
            e.svrg_init(z1 * 0 + z1, gamma, True)
            outs.append(z1)
            assert f1 < f0
        assert np.array_equal(outs[0], outs[1])


def test_full_scale_gradient_against_threaded_cpu_pass():
    """SURVEY.md §8d parity protocol at BASELINE.json's full size: the N = 2^22 × 4096 full-gradient pass (137.6 GB of row
    records) against a CPU pass over the same rows, regenerated on the fly on all host cores with long-double accumulation
    (oracle.full_gradient_synth; the matrix does not fit host memory).  SVRG_basic.jl:58-63, 88-92."""
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    if free < 150e9:
        pytest.skip("needs a 180 GB B200")
    N, d, seed = 1 << 22, 4096, 0x5EED0003
    x = np.random.default_rng(7).standard_normal(d) * 1e-3
    with Engine(0) as e:
        e.gen_synthetic(L.SYNTH_LASSO, N, d, seed, scale=float(N))
        g_dev = e.full_gradient(x, 1.0 / N)
        f_dev = e.objective(x)[0]
    g_cpu, f_sum = orc.full_gradient_synth(orc.SYN_LASSO, N, d, seed, float(N), x, 1.0 / N)
    assert rel(g_dev, g_cpu) <= 1e-12
    assert np.abs(g_dev - g_cpu).max() <= 1e-11 * np.abs(g_cpu).max()
    assert abs(f_dev - f_sum / N) <= 1e-12 * abs(f_sum / N)


def test_full_scale_table_invariants_c2_c5():
    """BASELINE configs 2 and 5 at full size (N = 2^20 × 1024 logistic; 2^18 blocks × 1024): the running averages the
    sequential kernels keep in registers must equal what the N×d tables hold in HBM — size-independent invariants of
    the reference's updates (SAGA_basic.jl:47,62: av = Σ s_i / N;  Finito_basic.jl:83,115: av = γ̂ Σ s_i/γ_i with equal
    γ_i;  ProShI_basic.jl:83,113-119: av = Σ s_i), checked after random steps with the table hazard path active."""
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    if free < 40e9:
        pytest.skip("needs 40 GB of HBM")
    N, d = 1 << 20, 1024
    with Engine(0) as e:
        e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0)
        e.set_reg(L.REG_NORML1, 1.0 / N)
        Lmax = 0.25 * e.max_row_sqnorm()
        idx = HostRNG(0x1D0002).rand_vec(N, N // 2)
        idx[1000:1040] = idx[999]                       # repeats inside the prefetch window
        e.saga_init(np.ones(d), 1 / (3 * Lmax), False)
        e.saga_steps(idx)
        av = e.get_vec(L.VEC_AV)
        assert rel(av * N, e.table_colsum()) < 1e-9
        z1 = e.get_vec(L.VEC_Z)
        e.saga_init(np.ones(d), 1 / (3 * Lmax), False)
        e.saga_steps(idx)
        assert np.array_equal(z1, e.get_vec(L.VEC_Z))   # bitwise reproducible at full size
        gam = np.full(N, 0.999 * N / Lmax)
        hat = 1 / np.sum(1 / gam)
        e.finito_init(np.ones(d), gam, hat)
        e.finito_steps(idx, np.arange(len(idx) + 1, dtype=np.int64))
        assert rel(e.get_vec(L.VEC_AV), e.table_colsum() * (hat / gam[0])) < 1e-9
        f0, f1 = sum(e.objective(np.ones(d))), sum(e.objective(e.get_vec(L.VEC_Z)))
        assert f1 < f0
    N, n = 1 << 18, 1024
    with Engine(0) as e:
        e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005)
        e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
        gam = 0.999 * N / np.full(N, 10.0 + 10.0 * N)
        e.proshi_init(np.zeros(n), gam, float(np.sum(gam)))
        idx = HostRNG(0x1D0005).rand_vec(N, N)
        idx[500:520] = idx[499]
        e.proshi_steps(idx, np.arange(N + 1, dtype=np.int64))                     # batch 1: producer-lane kernel
        assert rel(e.get_vec(L.VEC_AV), e.table_colsum()) < 1e-9
        sw = BatchSweeper(N, 4096, 2, HostRNG(1))
        bidx, bp = csr(sw.take(sw.d))
        e.proshi_steps(bidx, bp)                                                   # batch 4096: four-lanes-per-block kernel
        assert rel(e.get_vec(L.VEC_AV), e.table_colsum()) < 1e-9
        z = e.get_vec(L.VEC_Z)
        assert np.all(np.isfinite(z))


def test_soft_threshold_bits_match_the_reference_formula():
    """prox_NormL1 on the device is assembled from |x| − γλ and integer selects; it must give the bits of the reference's
    x + (x ≤ −γλ ? γλ : (x ≥ γλ ? −γλ : −x)) (ProximalOperators NormL1, restated in the oracle) — including +0 at and
    inside the threshold.  Exercised through SAGA's z = prox_g((1 − γ)·x0, γ) (SAGA_basic.jl:48) with γ = 1/2, λ = 2."""
    d = 16
    x0 = np.array([2.0, -2.0, 0.0, -0.0, 1.0, -1.0, 7.0, -7.0, 2.0 + 2 ** -50, -2.0 - 2 ** -50, 2.0 - 2 ** -51, -2.0 + 2 ** -51,
                   1e-310, -1e-310, 1e300, -1e300])
    rs = np.random.RandomState(0)
    A, b = rs.randn(4, d), rs.randn(4)
    p = orc.Problem(orc.LOSS_LS, A, b, np.ones(4)).set_reg(orc.REG_NORML1, lam=2.0)
    want = p.prox(0.5 * x0, 0.5)
    with Engine(0) as e:
        e.set_rows(L.LOSS_LS, A, b, 1.0)
        e.set_reg(L.REG_NORML1, 2.0)
        e.saga_init(x0, 0.5, False)
        got = e.get_vec(L.VEC_Z)
    assert np.array_equal(got.view(np.int64), want.view(np.int64)), (got, want)


def test_medium_scale_end_to_end_parity():
    """SURVEY.md §8d parity protocol at reduced scale: same inputs and the same index arrays into oracle and engine,
    N = 65 536 × 1024 logistic (the C2 generator) and N = 32 768 × 4096 Lasso (the C3 generator, 1 GiB); tolerances of the
    north star after K epochs: rel. objective ≤ 1e-8, ‖x − x_ref‖/‖x_ref‖ ≤ 1e-6."""
    # C2 shape: SAGA and Finito, one epoch of random single-sample steps each
    N, d = 1 << 16, 1024
    p, e = make_rows(orc.LOSS_LOGISTIC, N, d, 0x5EED0002, lam_reg=1.0 / N)
    Lmax = 0.25 * p.max_row_sqnorm()
    x0 = np.ones(d)
    idx = HostRNG(0x1D0002).rand_vec(N, N)
    ref = orc.SAGAState(p, x0, 1 / (3 * Lmax))
    ref.steps(idx)
    e.saga_init(x0, 1 / (3 * Lmax), False)
    e.saga_steps(idx)
    z = e.get_vec(L.VEC_Z)
    assert rel(z, ref.z) < REL_ITERATE
    assert abs(sum(e.objective(z)) - sum(p.objective(ref.z))) <= REL_OBJECTIVE * abs(sum(p.objective(ref.z)))
    gam = np.full(N, 0.999 * N / Lmax)
    reff = orc.FinitoState(p, x0, gam)
    reff.steps([idx[k:k + 1] for k in range(N // 2)])
    e.finito_init(x0, gam, reff.hat_gamma)
    e.finito_steps(idx[:N // 2], np.arange(N // 2 + 1, dtype=np.int64))
    z = e.get_vec(L.VEC_Z)
    assert rel(z, reff.z) < REL_ITERATE
    assert abs(sum(e.objective(z)) - sum(p.objective(reff.z))) <= REL_OBJECTIVE * abs(sum(p.objective(reff.z)))
    e.close()
    # C3 shape: SVRG++ with the bench's schedule m = N/16·2^k, three outer iterations
    N, d = 1 << 15, 4096
    p, e = make_rows(orc.LOSS_LS, N, d, 0x5EED0003, lam_reg=N / 100.0)
    gamma = 1 / (7 * N * p.max_row_sqnorm())
    ref = orc.SVRGState(p, np.zeros(d), gamma, m=N // 16, plus=True)
    e.svrg_init(np.zeros(d), gamma, True)
    rng, m = HostRNG(0x1D0003), N // 16
    for _ in range(3):
        idx = rng.rand_vec(N, m)
        ref.epoch(idx)
        e.svrg_epoch(idx)
        m *= 2
    x = e.get_vec(L.VEC_Z_FULL)
    assert rel(x, ref.z_full) < REL_ITERATE
    f, f_ref = sum(e.objective(x)), sum(p.objective(ref.z_full))
    assert abs(f - f_ref) <= REL_OBJECTIVE * abs(f_ref) and f_ref < sum(p.objective(np.zeros(d)))
    e.close()


def test_large_d_sequential_kernels():
    """d = 8192: 1024 columns per CTA, 8 columns per thread; the SAGA/Finito ring with table slices no longer fits in shared
    memory, so the table rows take the register-prefetch path on their own (no environment knob)."""
    N, d = 96, 8192
    p, e = make_rows(orc.LOSS_LS, N, d, 0xB16D, lam_reg=0.05)
    Lmax = N * p.max_row_sqnorm()
    x0 = np.full(d, 0.01)
    idx = HostRNG(3).rand_vec(N, 3 * N)
    ref = orc.SVRGState(p, x0, 1 / (7 * Lmax), m=N, plus=False)
    e.svrg_init(x0, 1 / (7 * Lmax), False)
    for k in range(3):
        ref.epoch(idx[k * N:(k + 1) * N])
        e.svrg_epoch(idx[k * N:(k + 1) * N])
    assert rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-9
    refs = orc.SAGAState(p, x0, 1 / (3 * Lmax))
    refs.steps(idx)
    e.saga_init(x0, 1 / (3 * Lmax), False)
    e.saga_steps(idx)
    assert rel(e.get_vec(L.VEC_Z), refs.z) < 1e-9 and rel(e.get_table_rows(), refs.s) < 1e-9
    gam = 0.999 * N / (N * np.sum(p.A * p.A, axis=1))
    reff = orc.FinitoState(p, x0, gam)
    reff.steps([idx[k:k + 1] for k in range(2 * N)])
    e.finito_init(x0, gam, reff.hat_gamma)
    e.finito_steps(idx[:2 * N], np.arange(2 * N + 1, dtype=np.int64))
    assert rel(e.get_vec(L.VEC_Z), reff.z) < 1e-9 and rel(e.get_table_rows(), reff.s) < 1e-9
    e.close()


# ----------------------------------------------------------------------------
# "the iterator state is the checkpoint" (SAGA_basic.jl:11-20): read the state out, put it back into a fresh context,
# continue — bit for bit the uninterrupted run for the table solvers
@pytest.mark.parametrize("alg", ["saga", "finito", "proshi"])
def test_checkpoint_restore_continues_bitwise(alg):
    N, d = 300, 256
    rng = HostRNG(21)
    if alg == "proshi":
        def fresh():
            e = Engine(0)
            e.gen_synthetic(L.SYNTH_SHARING, N, d, 0x5EED0005)
            e.set_reg(L.REG_INDBOX, -np.inf, np.ones(d))
            return e
        gam = 0.999 * N / np.full(N, 10.0 + 10.0 * N)
        hat = float(np.sum(gam))
        x0 = np.zeros(d)
    else:
        def fresh():
            e = Engine(0)
            e.gen_synthetic(L.SYNTH_LASSO, N, d, 0xC0FFEE, scale=float(N))
            e.set_reg(L.REG_NORML1, 0.05)
            return e
        x0 = np.full(d, 0.1)
    a = fresh()
    Lmax = N * a.max_row_sqnorm() if alg != "proshi" else None
    if alg == "finito":
        gam = np.linspace(0.6, 1.0, N) * (0.999 * N / Lmax)
        hat = 1 / np.sum(1 / gam)
    idx1, idx2 = rng.rand_vec(N, 2 * N), rng.rand_vec(N, 2 * N)
    bp = np.arange(2 * N + 1, dtype=np.int64)

    def steps(e, idx):
        if alg == "saga":
            e.saga_steps(idx)
        elif alg == "finito":
            e.finito_steps(idx, bp)
        else:
            e.proshi_steps(idx, bp)

    if alg == "saga":
        a.saga_init(x0, 1 / (3 * Lmax), False)
    elif alg == "finito":
        a.finito_init(x0, gam, hat)
    else:
        a.proshi_init(x0, gam, hat)
    steps(a, idx1)
    ck = {"z": a.get_vec(L.VEC_Z), "av": a.get_vec(L.VEC_AV), "s": a.get_table_rows()}      # the checkpoint
    steps(a, idx2)
    b = fresh()
    if alg == "saga":
        b.solver_restore(2, gamma=1 / (3 * Lmax))
    else:
        b.solver_restore(3 if alg == "finito" else 5, gamma_N=gam, hat_gamma=hat)
    b.set_vec(L.VEC_Z, ck["z"])
    b.set_vec(L.VEC_AV, ck["av"])
    b.set_table_rows(ck["s"][:100])
    b.set_table_rows(ck["s"][100:], 100)
    steps(b, idx2)
    assert np.array_equal(a.get_vec(L.VEC_Z), b.get_vec(L.VEC_Z))
    assert np.array_equal(a.get_vec(L.VEC_AV), b.get_vec(L.VEC_AV))
    assert np.array_equal(a.get_table_rows(), b.get_table_rows())
    a.close()
    b.close()


def test_checkpoint_restore_svrg():
    N, d = 512, 256
    def fresh():
        e = Engine(0)
        e.gen_synthetic(L.SYNTH_LASSO, N, d, 0xC0FFEE, scale=float(N))
        e.set_reg(L.REG_NORML1, 0.05)
        return e
    a = fresh()
    gamma = 1 / (7 * N * a.max_row_sqnorm())
    rng = HostRNG(5)
    i1, i2 = rng.rand_vec(N, N // 2), rng.rand_vec(N, N)
    a.svrg_init(np.zeros(d), gamma, True)
    a.svrg_epoch(i1)
    ck = {k: a.get_vec(v) for k, v in (("z", L.VEC_Z), ("zf", L.VEC_Z_FULL), ("w", L.VEC_W), ("av", L.VEC_AV))}
    a.svrg_epoch(i2)
    b = fresh()
    b.solver_restore(1, gamma=gamma, flag=True)
    for k, v in (("z", L.VEC_Z), ("zf", L.VEC_Z_FULL), ("w", L.VEC_W), ("av", L.VEC_AV)):
        b.set_vec(v, ck[k])
    b.svrg_epoch(i2)
    # the restored context recomputes c_i(z_full) inside the inner kernel instead of taking it from the last pass: rounding-level
    assert rel(b.get_vec(L.VEC_Z_FULL), a.get_vec(L.VEC_Z_FULL)) < 1e-12
    assert rel(b.get_vec(L.VEC_AV), a.get_vec(L.VEC_AV)) < 1e-10
    a.close()
    b.close()


def test_objective_indbox_is_infinite_outside_the_box():
    N, d = 64, 32
    with Engine(0) as e:
        e.gen_synthetic(L.SYNTH_LASSO, N, d, 3, scale=float(N))
        e.set_reg(L.REG_INDBOX, -1.0, 1.0)
        assert e.objective(np.full(d, 0.5))[1] == 0.0
        x = np.full(d, 0.5)
        x[7] = 1.0 + 1e-12
        assert e.objective(x)[1] == np.inf
        e.set_reg(L.REG_INDBOX, -np.inf, np.linspace(0.0, 1.0, d))
        assert e.objective(np.zeros(d))[1] == 0.0 and e.objective(np.full(d, 0.5))[1] == np.inf


def test_out_of_range_index_leaves_the_state_untouched():
    """The reference throws BoundsError at F[i] before any state changes (SAGA_basic.jl:56): an invalid index anywhere in a call
    means none of the call's steps run; the error surfaces at the next synchronising call and the context stays usable."""
    N, d = 200, 128
    with Engine(0) as e:
        e.gen_synthetic(L.SYNTH_LASSO, N, d, 11, scale=float(N))
        e.set_reg(L.REG_NORML1, 0.05)
        Lmax = N * e.max_row_sqnorm()
        e.saga_init(np.full(d, 0.2), 1 / (3 * Lmax), False)
        good = HostRNG(1).rand_vec(N, 50)
        e.saga_steps(good)
        z, av, s = e.get_vec(L.VEC_Z), e.get_vec(L.VEC_AV), e.get_table_rows()
        bad = good.copy()
        bad[30] = N + 1
        e.saga_steps(bad)
        with pytest.raises(CiaoError) as ei:
            e.get_vec(L.VEC_Z)
        assert ei.value.code == -1
        assert np.array_equal(e.get_vec(L.VEC_Z), z) and np.array_equal(e.get_vec(L.VEC_AV), av)
        assert np.array_equal(e.get_table_rows(), s)
        e.saga_steps(good)                                   # still usable
        assert not np.array_equal(e.get_vec(L.VEC_Z), z)


def test_finito_adaptive_random_restart():
    """Finito_adaptive.jl:77-83 through the C ABI: the library calls back into the host for the perturbation draws of the
    degenerate components, in the reference's order; same draws → same γ_i, γ̂, av, z as the oracle, and the steps continue from there."""
    from ciaoalgorithms_jl_b200.sampling import JuliaRNG
    N, d = 40, 96
    rs = np.random.default_rng(5)
    A = rs.standard_normal((N, d))
    for i in (3, 17, 18):                       # small integers with Σ_k a_ik = 0: ∇f_i(x0 + 1) == ∇f_i(x0) exactly, in any order
        A[i, :d // 2] = rs.integers(-3, 4, size=d // 2)
        A[i, d // 2:] = -A[i, :d // 2]
    b = rs.standard_normal(N)
    p = orc.Problem(orc.LOSS_LS, A, b, np.full(N, float(N))).set_reg(orc.REG_NORML1, lam=0.05)
    x0 = np.zeros(d)
    with Engine(0) as e:
        e.set_rows(L.LOSS_LS, A, b, float(N))
        e.set_reg(L.REG_NORML1, 0.05)
        with pytest.raises(CiaoError) as ei:
            e.finito_adaptive_init(x0)
        assert ei.value.code == -4
        ra, rb, calls = JuliaRNG(11), JuliaRNG(11), []

        def pert(i, t):
            calls.append((i, t))
            return rb.rand_pm(t, d)

        ref = orc.FinitoAdaptiveState(p, x0, perturb=lambda i, t: ra.rand_pm(t, d))
        e.finito_adaptive_init(x0, perturb=pert)
        assert [c[0] for c in calls if c[1] == 1] == [4, 18, 19]
        gam, fi_x, _, hat, _ = e.finito_adaptive_get(fi_x=True)
        assert rel(gam, ref.gamma) < 1e-12 and abs(hat - ref.hat_gamma) <= 1e-12 * ref.hat_gamma
        assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-10 and rel(e.get_vec(L.VEC_Z), ref.z) < 1e-10
        idx = AdaptiveSweeper(N, 1, HostRNG(2)).take(6 * N)
        assert ref.steps(idx) == e.finito_adaptive_steps(idx)
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
        assert rel(e.finito_adaptive_get()[0], ref.gamma) < 1e-12


def test_interleaved_shards_host_rows_generator_and_whole_problem_agree():
    """ciao_set_row_interleave: rank k of G holds the blocks of B rows number k, k + G, …  Rows handed over from the host in that
    order and rows generated on the device must give the same shard, and the shards' partial gradients must add up to the whole
    problem's (no communicator here: every "rank" is a context of its own on this GPU)."""
    from ciaoalgorithms_jl_b200.sampling import interleaved_rows
    N, d, B, G, seed = 500, 40, 16, 3, 0x11AB
    A, rhs = orc.gen_rows(orc.SYN_LASSO, d, seed, 0, N)
    x = np.random.default_rng(3).standard_normal(d)
    with Engine(0) as full:
        full.gen_synthetic(L.SYNTH_LASSO, N, d, seed, scale=float(N))
        g_full = full.full_gradient(x, 1.0)
    total = np.zeros(d)
    for k in range(G):
        gl = interleaved_rows(N, B, G, k)
        with Engine(0) as a, Engine(0) as b:
            a.set_row_interleave(B, k, G)
            a.set_rows(L.LOSS_LS, A[gl], rhs[gl], float(N), N_total=N, row0=k * B)
            b.set_row_interleave(B, k, G)
            b.gen_synthetic(L.SYNTH_LASSO, N, d, seed, scale=float(N), row0=k * B, n_rows=len(gl))
            ga, gb = a.full_gradient(x, 1.0), b.full_gradient(x, 1.0)
            assert np.array_equal(ga, gb)
            total += ga
    assert rel(total, g_full) < 1e-13
    with Engine(0) as e:
        e.set_row_interleave(B, 1, G)
        with pytest.raises(Exception, match="interleaved shard"):
            e.gen_synthetic(L.SYNTH_LASSO, N, d, seed, scale=float(N), row0=B, n_rows=N // G)     # rank 1 owns 164 rows, not 166

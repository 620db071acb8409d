"""M×d block components and genuinely complex data through the general block kernel (csrc/blockseq.cu),
against the oracle's restatement of ProximalOperators' dense LeastSquares / Precompose(LogisticLoss) with an M×d matrix and of the
complex NormL1 prox.  Also: the block kernel on ordinary M = 1 problems (CIAO_FORCE_BLOCK_KERNEL=1) — a second, independently
written CUDA path for every sequential loop, which must agree with the oracle exactly like the tuned cluster kernels do."""
import numpy as np
import pytest

from oracle import oracle as orc
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200 import operators as ops
from ciaoalgorithms_jl_b200 import solvers as S
from ciaoalgorithms_jl_b200.engine import CiaoError, Engine
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, LFinitoSweeper, csr

gpu = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def run_all_loops(p, e, N, d, Lmax, gam, reg_pairs=False):
    """the same SVRG / SAGA / SAG / Finito / LFinito calls on the oracle problem p and the engine e; asserts parity after each"""
    x0 = np.full(d, 0.1)
    rng = HostRNG(4)
    x = np.random.default_rng(1).standard_normal(d) * 0.3
    assert rel(e.full_gradient(x, 1.0 / N), p.full_gradient(x, 1.0 / N)) < 1e-11
    f_e, g_e = e.objective(x)
    f_o, g_o = p.objective(x)
    assert abs(f_e - f_o) <= 1e-11 * abs(f_o) and abs(g_e - g_o) <= 1e-12 * max(1.0, abs(g_o))
    # SVRG and SVRG++
    for plus in (False, True):
        ref = orc.SVRGState(p, x0, 1 / (7 * Lmax), m=N, plus=plus)
        e.svrg_init(x0, 1 / (7 * Lmax), plus)
        assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-11
        for _ in range(3):
            idx = rng.rand_vec(N, ref.m)          # m doubles after every epoch of SVRG++ (SVRG_basic.jl:93)
            ref.epoch(idx)
            e.svrg_epoch(idx)
        assert rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-9 and rel(e.get_vec(L.VEC_W), ref.w) < 1e-9
        assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    # SAGA / SAG
    for sag in (False, True):
        ref = orc.SAGAState(p, x0, 1 / ((16 if sag else 3) * Lmax), sag=sag)
        e.saga_init(x0, 1 / ((16 if sag else 3) * Lmax), sag)
        assert rel(e.get_table_rows(), ref.s) < 1e-11 and rel(e.get_vec(L.VEC_AV), ref.av) < 1e-11
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-11
        idx = rng.rand_vec(N, 5 * N)
        idx[7:10] = idx[6]
        ref.steps(idx)
        e.saga_steps(idx)
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9 and rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
        assert rel(e.get_table_rows(), ref.s) < 1e-9
    # Finito: three sweepings, batches
    for sweeping, batch in ((1, 1), (2, 3), (3, 2)):
        ref = orc.FinitoState(p, x0, gam)
        e.finito_init(x0, gam, ref.hat_gamma)
        assert rel(e.get_table_rows(), ref.s) < 1e-11 and rel(e.get_vec(L.VEC_AV), ref.av) < 1e-10
        batches = BatchSweeper(N, batch, sweeping, rng).take(3 * (-(-N // batch)))
        ref.steps(batches)
        e.finito_steps(*csr(batches))
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9 and rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
        assert rel(e.get_table_rows(), ref.s) < 1e-9
    # LFinito
    for sweeping, batch in ((2, 1), (3, 4)):
        ref = orc.LFinitoState(p, x0, gam, batch=batch)
        e.lfinito_init(x0, gam, ref.hat_gamma)
        sw = LFinitoSweeper(N, batch, sweeping, rng)
        for _ in range(2):
            o = sw.next()
            ref.outer(o)
            e.lfinito_outer(o, batch)
        assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9 and rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-9
        assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8


@gpu
@pytest.mark.parametrize("kind", [orc.LOSS_LS, orc.LOSS_LOGISTIC])
@pytest.mark.parametrize("N,d", [(40, 96), (25, 2100)])
def test_block_kernel_on_ordinary_rows(kind, N, d, monkeypatch):
    monkeypatch.setenv("CIAO_FORCE_BLOCK_KERNEL", "1")            # read by ciao_create
    syn = orc.SYN_LASSO if kind == orc.LOSS_LS else orc.SYN_LOGISTIC
    A, rhs = orc.gen_rows(syn, d, 0xB10C + d, 0, N)
    sc = float(N) if kind == orc.LOSS_LS else 1.0
    lam = 0.05 if kind == orc.LOSS_LS else 1.0 / N
    p = orc.Problem(kind, A, rhs, np.full(N, sc)).set_reg(orc.REG_NORML1, lam=lam)
    Li = np.sum(A * A, axis=1) * (N if kind == orc.LOSS_LS else 0.25)
    with Engine(0) as e:
        e.set_rows(kind, A, rhs, sc)
        e.set_reg(L.REG_NORML1, lam)
        run_all_loops(p, e, N, d, Li.max(), 0.999 * N / Li)


@gpu
@pytest.mark.parametrize("kind", [orc.LOSS_LS, orc.LOSS_LOGISTIC])
@pytest.mark.parametrize("N,M,d", [(30, 3, 96), (12, 7, 130), (9, 2, 2100)])
def test_row_block_components_all_loops(kind, N, M, d):
    rs = np.random.default_rng(N * M + d)
    A = rs.standard_normal((N * M, d)) / np.sqrt(d)
    if kind == orc.LOSS_LS:
        rhs, sc, lam = rs.standard_normal(N * M), np.full(N, float(N)), 0.05
    else:
        rhs, sc, lam = np.sign(rs.standard_normal(N * M)), np.ones(N), 1.0 / N
    p = orc.Problem(kind, A, rhs, sc, rows_per_component=M).set_reg(orc.REG_NORML1, lam=lam)
    Li = np.array([np.linalg.norm(A[i * M:(i + 1) * M], 2) ** 2 for i in range(N)]) * (N if kind == orc.LOSS_LS else 0.25)
    with Engine(0) as e:
        e.set_row_blocks(kind, A, rhs, M, sc)
        e.set_reg(L.REG_NORML1, lam)
        run_all_loops(p, e, N, d, Li.max(), 0.999 * N / Li)
        with pytest.raises(CiaoError) as ei:                      # adaptive Finito: not for block components
            e.finito_adaptive_init(np.zeros(d))
        assert ei.value.code == -4


@gpu
def test_genuinely_complex_lasso_through_the_realified_blocks():
    """test_lasso.jl:3 with data that really leaves the real axis: complex rows a_i, b_i, x0; NormL1 on complex numbers."""
    N, dc = 24, 20
    rs = np.random.default_rng(3)
    A = (rs.standard_normal((N, dc)) + 1j * rs.standard_normal((N, dc))) / np.sqrt(dc)
    b = rs.standard_normal(N) + 1j * rs.standard_normal(N)
    x0 = 0.1 * (rs.standard_normal(dc) + 1j * rs.standard_normal(dc))
    Ar = np.vstack([L.realify_rows(A[i:i + 1]) for i in range(N)])
    br = np.concatenate([L.realify_vec(b[i:i + 1]) for i in range(N)])
    p = orc.Problem(orc.LOSS_LS, Ar, br, np.full(N, float(N)), rows_per_component=2).set_reg(orc.REG_NORML1_PAIRS, lam=0.05)
    Li = N * np.sum(np.abs(A) ** 2, axis=1)
    with Engine(0) as e:
        e.set_row_blocks(L.LOSS_LS, Ar, br, 2, float(N))
        e.set_reg(L.REG_NORML1_PAIRS, 0.05)
        run_all_loops(p, e, N, 2 * dc, Li.max(), 0.999 * N / Li)
    # through the reference-shaped API: complex F, complex x0, complex solution; same indices → the oracle's iterate
    F = [ops.LeastSquares(A[i:i + 1, :], b[i:i + 1], float(N)) for i in range(N)]
    g = ops.NormL1(0.05)
    from ciaoalgorithms_jl_b200.sampling import JuliaRNG      # its vector draw is by construction the sequence of scalar draws
    x, it = S.SAGA(gamma=1 / (3 * Li.max()), maxit=400)(x0, F=F, g=g, N=N, rng=JuliaRNG(8))
    ref = orc.SAGAState(p, L.realify_vec(x0), 1 / (3 * Li.max()))
    ref.steps(JuliaRNG(8).rand_vec(N, 399))
    assert x.dtype == np.complex128 and it == 400
    assert rel(L.realify_vec(x), ref.z) < 1e-9
    cost = lambda v: 0.5 * np.linalg.norm(A @ v - b) ** 2 + 0.05 * np.sum(np.abs(v))     # noqa: E731
    assert cost(x) < cost(x0)
    x2, _ = S.SVRG(gamma=1 / (7 * Li.max()), maxit=30)(x0, F=F, g=g, N=N, rng=HostRNG(8))
    x3, _ = S.Finito(maxit=400, sweeping=2)(x0, F=F, g=g, L=Li, N=N, rng=HostRNG(8))
    assert x2.dtype == np.complex128 and cost(x2) < cost(x0) and cost(x3) < cost(x0)


def test_realification_is_the_complex_arithmetic():
    """CPU: the real block [Re; Im] of a complex row reproduces A·x and Aᴴ·res of complex arithmetic, and the oracle's block
    LeastSquares + pair soft-threshold reproduce complex SAGA steps written directly with numpy complex numbers."""
    rs = np.random.default_rng(0)
    N, dc = 7, 5
    A = rs.standard_normal((N, dc)) + 1j * rs.standard_normal((N, dc))
    b = rs.standard_normal(N) + 1j * rs.standard_normal(N)
    x = rs.standard_normal(dc) + 1j * rs.standard_normal(dc)
    R = L.realify_rows(A[2:3])
    u = R @ L.realify_vec(x)
    assert np.allclose(u[0] + 1j * u[1], A[2] @ x, rtol=0, atol=1e-14)
    res = A[2] @ x - b[2]
    assert np.allclose(L.complexify_vec(R.T @ np.array([res.real, res.imag])), np.conj(A[2]) * res, rtol=0, atol=1e-14)
    # complex SAGA, directly: f_i = (N/2)|a_i·x − b_i|², ∇f_i = N·conj(a_i)(a_i·x − b_i), prox = complex soft-threshold
    lam, gamma = 0.3, 0.01
    soft = lambda v, t: np.where(np.abs(v) > 0, v / np.maximum(np.abs(v), 1e-300) * np.maximum(np.abs(v) - t, 0), 0)   # noqa: E731
    grad = lambda i, v: N * np.conj(A[i]) * (A[i] @ v - b[i])      # noqa: E731
    s = [grad(i, x) for i in range(N)]
    av = sum(s) / N
    z = soft((1 - gamma) * x, gamma * lam)
    idx = rs.integers(1, N + 1, size=60)
    for i1 in idx:
        i = i1 - 1
        gi = grad(i, z)
        w = z - gamma * (gi - s[i] + av)
        av = av + (gi - s[i]) / N
        z = soft(w, gamma * lam)
        s[i] = gi
    Ar = np.vstack([L.realify_rows(A[i:i + 1]) for i in range(N)])
    br = np.concatenate([L.realify_vec(b[i:i + 1]) for i in range(N)])
    p = orc.Problem(orc.LOSS_LS, Ar, br, np.full(N, float(N)), rows_per_component=2).set_reg(orc.REG_NORML1_PAIRS, lam=lam)
    st = orc.SAGAState(p, L.realify_vec(x), gamma)
    st.steps(idx.astype(np.int64))
    assert np.allclose(L.complexify_vec(st.z), z, rtol=0, atol=1e-12)
    assert np.allclose(L.complexify_vec(st.av), av, rtol=0, atol=1e-11)


def test_oracle_row_blocks_are_the_dense_least_squares_gradient():
    rs = np.random.default_rng(1)
    N, M, d = 5, 3, 4
    A, b = rs.standard_normal((N * M, d)), rs.standard_normal(N * M)
    lam = rs.uniform(0.5, 2.0, N)
    p = orc.Problem(orc.LOSS_LS, A, b, lam, rows_per_component=M)
    x = rs.standard_normal(d)
    for i in range(N):
        Ai, bi = A[i * M:(i + 1) * M], b[i * M:(i + 1) * M]
        g, f = p.gradient(i, x)
        assert np.allclose(g, lam[i] * Ai.T @ (Ai @ x - bi), rtol=0, atol=1e-13)
        assert abs(f - lam[i] / 2 * np.sum((Ai @ x - bi) ** 2)) < 1e-13
    assert np.allclose(p.full_gradient(x), sum(lam[i] * A[i * M:(i + 1) * M].T @ (A[i * M:(i + 1) * M] @ x - b[i * M:(i + 1) * M]) for i in range(N)))

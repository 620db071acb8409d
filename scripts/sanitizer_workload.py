"""Every kernel of libciao_cuda once, on the reference's own tiny fixtures (N ≤ 8) and on small synthetic shapes, for
compute-sanitizer (memcheck / racecheck / synccheck / initcheck; scripts/sanitize.sh).  Results are checked against the oracle
as in the tests, so a sanitizer-clean run is also a correct one."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ciao_pkg; ciao_pkg.load()
import fixtures
from oracle import oracle as orc
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, AdaptiveSweeper, csr

rel = lambda a, b: np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)
small = "--small" in sys.argv


def rows(kind, N, d, seed):
    syn = orc.SYN_LASSO if kind == orc.LOSS_LS else orc.SYN_LOGISTIC
    A, rhs = orc.gen_rows(syn, d, seed, 0, N)
    sc = float(N) if kind == orc.LOSS_LS else 1.0
    p = orc.Problem(kind, A, rhs, np.full(N, sc)).set_reg(orc.REG_NORML1, lam=0.05 if kind == orc.LOSS_LS else 1.0 / N)
    e = Engine(0)
    e.gen_synthetic(L.SYNTH_LASSO if kind == orc.LOSS_LS else L.SYNTH_LOGISTIC, N, d, seed, scale=sc)
    e.set_reg(L.REG_NORML1, 0.05 if kind == orc.LOSS_LS else 1.0 / N)
    return p, e


shapes = [(orc.LOSS_LS, 6, 3), (orc.LOSS_LOGISTIC, 8, 5), (orc.LOSS_LS, 40, 256)] + ([] if small else [(orc.LOSS_LOGISTIC, 64, 1024)])
for kind, N, d in shapes:
    p, e = rows(kind, N, d, 0x5A + d)
    Lmax = p.max_row_sqnorm() * (N if kind == orc.LOSS_LS else 0.25)
    x0 = np.full(d, 0.1)
    rng = HostRNG(3)
    # passes + SVRG
    assert rel(e.full_gradient(x0, 1.0 / N), p.full_gradient(x0, 1.0 / N)) < 1e-11
    ref = orc.SVRGState(p, x0, 1 / (7 * Lmax), m=N, plus=False)
    e.svrg_init(x0, 1 / (7 * Lmax), False)
    for _ in range(2):
        idx = rng.rand_vec(N, 2 * N)
        ref.epoch(idx); e.svrg_epoch(idx)
    assert rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-9
    # SAGA (repeats inside the hazard window are certain at these N)
    ref = orc.SAGAState(p, x0, 1 / (3 * Lmax))
    e.saga_init(x0, 1 / (3 * Lmax), False)
    idx = rng.rand_vec(N, 6 * N)
    ref.steps(idx); e.saga_steps(idx)
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9 and rel(e.get_table_rows(), ref.s) < 1e-9
    # Finito with batches, LFinito
    gam = 0.999 * N / (np.sum(p.A * p.A, axis=1) * (N if kind == orc.LOSS_LS else 0.25))
    ref = orc.FinitoState(p, x0, gam)
    e.finito_init(x0, gam, ref.hat_gamma)
    batches = BatchSweeper(N, 2, 3, rng).take(3 * N)
    ref.steps(batches); e.finito_steps(*csr(batches))
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    refl = orc.LFinitoState(p, x0, gam, batch=2)
    e.lfinito_init(x0, gam, refl.hat_gamma)
    o = np.arange(1, refl.nb + 1, dtype=np.int64)
    refl.outer(o); e.lfinito_outer(o, 2)
    assert rel(e.get_vec(L.VEC_Z), refl.z) < 1e-9
    # adaptive Finito
    refa = orc.FinitoAdaptiveState(p, x0)
    e.finito_adaptive_init(x0)
    idx = AdaptiveSweeper(N, 1, rng).take(4 * N)
    assert refa.steps(idx) == e.finito_adaptive_steps(idx)
    assert rel(e.get_vec(L.VEC_Z), refa.z) < 1e-9
    e.close()
    print(f"rows kind={kind} N={N} d={d}: ok", flush=True)

# minibatch passes (≥ 256 rows per batch)
p, e = rows(orc.LOSS_LS, 600, 128, 77)
Lmax = p.max_row_sqnorm() * 600
gam = np.full(600, 0.999 * 600 / Lmax)
x0 = np.full(128, 0.1)
ref = orc.FinitoState(p, x0, gam)
e.finito_init(x0, gam, ref.hat_gamma)
batches = BatchSweeper(600, 256, 2, HostRNG(1)).take(3)
ref.steps(batches); e.finito_steps(*csr(batches))
assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
e.close()
print("minibatch: ok", flush=True)

# ProShI: the reference fixture and a synthetic sharing problem, batch 1 and minibatch
fx = fixtures.sharing()
for (N, n, use_fx) in ((3, 2, True), (70, 96, False)):
    if use_fx:
        Q, ql, box, eta, Li = fx["Qdiag"], fx["qlin"], fx["box"], fx["eta"], fx["L"]
    else:
        Q, _ = orc.gen_rows(orc.SYN_SHARING, n, 0x5EED0005, 0, N)
        ql, box, eta = np.ones((N, n)), (-2.0, 2.0), 10.0 * N
        Li = np.abs(Q).max(axis=1) + eta
    p = orc.Problem(orc.LOSS_DIAGQUAD, Q, ql, box=box, eta=eta).set_reg(orc.REG_INDBOX, lo=-np.inf, hi=np.ones(n))
    gam = 0.999 * N / Li
    ref = orc.ProshiState(p, np.zeros(n), gam)
    with Engine(0) as e:
        e.set_blocks(Q, ql, box, eta)
        e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
        e.proshi_init(np.zeros(n), gam, ref.hat_gamma)
        for r in ((1, 2) if use_fx else (1, 64)):
            batches = BatchSweeper(N, r, 2, HostRNG(2)).take(3 * (-(-N // r)))
            ref.steps(batches); e.proshi_steps(*csr(batches))
        assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-9
        out = np.empty((N, n))
        e.proshi_solution(out)
        assert rel(out, ref.solution()) < 1e-9
    print(f"sharing N={N} n={n}: ok", flush=True)
print("SANITIZER_WORKLOAD_OK")

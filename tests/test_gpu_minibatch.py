"""Minibatch steps of Finito / LFinito with large static batches run as streaming passes
(batch.cu) — parity against the oracle's sequential loop."""
import numpy as np
import pytest

from oracle import oracle as orc
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, LFinitoSweeper, csr
from test_gpu_parity import make_rows, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,N,d,batch", [(orc.LOSS_LS, 1500, 64, 256), (orc.LOSS_LOGISTIC, 2100, 1024, 512),
                                             (orc.LOSS_LS, 1024, 4096, 256), (orc.LOSS_LOGISTIC, 900, 130, 300),
                                             (orc.LOSS_LS, 6000, 64, 3000)])   # the last: several sub-groups of 32 threads per SM
@pytest.mark.parametrize("sweeping", [2, 3])
@pytest.mark.parametrize("launch", ["persistent", "per_batch"])
def test_finito_static_minibatch_pass(kind, N, d, batch, sweeping, launch, monkeypatch):
    monkeypatch.setenv("CIAO_BATCH_PER_LAUNCH", "1" if launch == "per_batch" else "0")   # read by ciao_create
    p, e = make_rows(kind, N, d, 0xBA7 + d, lam_reg=0.05 if kind == orc.LOSS_LS else 1.0 / N)
    Li = np.sum(p.A * p.A, axis=1) * (N if kind == orc.LOSS_LS else 0.25)
    gam = 0.999 * N / Li
    x0 = np.full(d, 0.1)
    ref = orc.FinitoState(p, x0, gam)
    e.finito_init(x0, gam, ref.hat_gamma)
    batches = BatchSweeper(N, batch, sweeping, HostRNG(4)).take(3 * (-(-N // batch)) + 1)
    ref.steps(batches)
    idx, bp = csr(batches)
    e.finito_steps(idx, bp)
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    assert rel(e.get_table_rows(), ref.s) < 1e-9
    z1 = e.get_vec(L.VEC_Z)
    e.finito_init(x0, gam, ref.hat_gamma)
    e.finito_steps(idx, bp)
    assert np.array_equal(z1, e.get_vec(L.VEC_Z))                  # bitwise reproducible
    e.close()


@pytest.mark.parametrize("kind,N,d,batch", [(orc.LOSS_LS, 1500, 64, 256), (orc.LOSS_LOGISTIC, 2100, 1024, 700), (orc.LOSS_LS, 6000, 64, 3000)])
@pytest.mark.parametrize("sweeping", [2, 3])
@pytest.mark.parametrize("launch", ["persistent", "per_batch"])
def test_lfinito_minibatch_sweep_pass(kind, N, d, batch, sweeping, launch, monkeypatch):
    monkeypatch.setenv("CIAO_BATCH_PER_LAUNCH", "1" if launch == "per_batch" else "0")
    p, e = make_rows(kind, N, d, 0xBA8 + d, lam_reg=0.05 if kind == orc.LOSS_LS else 1.0 / N)
    Li = np.sum(p.A * p.A, axis=1) * (N if kind == orc.LOSS_LS else 0.25)
    gam = 0.999 * N / Li
    x0 = np.full(d, 0.1)
    ref = orc.LFinitoState(p, x0, gam, batch)
    e.lfinito_init(x0, gam, ref.hat_gamma)
    sw = LFinitoSweeper(N, batch, sweeping, HostRNG(2))
    for _ in range(3):
        order = sw.next()
        ref.outer(order)
        e.lfinito_outer(order, batch)
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    assert rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    e.close()


@pytest.mark.parametrize("kind,N,d,batch", [(orc.LOSS_LOGISTIC, 2100, 1024, 512), (orc.LOSS_LS, 1500, 64, 256), (orc.LOSS_LS, 1024, 4096, 256),
                                             (orc.LOSS_LOGISTIC, 4500, 1024, 2200), (orc.LOSS_LS, 4500, 256, 2100)])
@pytest.mark.parametrize("exchange", ["words", "barrier"])
def test_finito_minibatch_one_epoch_per_call(kind, N, d, batch, exchange, monkeypatch):
    """One epoch per call (the windows of a call are pairwise disjoint; the epoch counter of the exchange keeps counting across the
    launches).  Three calls = three epochs against the oracle's sequential loop."""
    monkeypatch.setenv("CIAO_BATCH_EXCHANGE", exchange)   # read at every call
    p, e = make_rows(kind, N, d, 0xBA9 + d, lam_reg=0.05 if kind == orc.LOSS_LS else 1.0 / N)
    Li = np.sum(p.A * p.A, axis=1) * (N if kind == orc.LOSS_LS else 0.25)
    gam = 0.999 * N / Li
    x0 = np.full(d, 0.1)
    ref = orc.FinitoState(p, x0, gam)
    e.finito_init(x0, gam, ref.hat_gamma)
    sw = BatchSweeper(N, batch, 3, HostRNG(7))
    for _ in range(3):
        batches = sw.take(sw.d)
        ref.steps(batches)
        e.finito_steps(*csr(batches))
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    assert rel(e.get_table_rows(), ref.s) < 1e-9
    e.close()


@pytest.mark.parametrize("exchange", ["words", "barrier"])
def test_finito_minibatch_rows_change_window(exchange, monkeypatch):
    """Contiguous batches whose boundaries differ from pass to pass inside ONE call: a row is then handled by different CTAs,
    and its table row written in one batch is read by another CTA a few batches later (the fenced variant of the exchange)."""
    monkeypatch.setenv("CIAO_BATCH_EXCHANGE", exchange)
    kind, N, d = orc.LOSS_LOGISTIC, 2304, 1024
    p, e = make_rows(kind, N, d, 0xBAA, lam_reg=1.0 / N)
    gam = 0.999 * N / (np.sum(p.A * p.A, axis=1) * 0.25)
    x0 = np.full(d, 0.1)
    ref = orc.FinitoState(p, x0, gam)
    e.finito_init(x0, gam, ref.hat_gamma)
    batches = []
    for width in (256, 384, 288, 576, 256):
        batches += [np.arange(lo + 1, min(lo + width, N) + 1, dtype=np.int64) for lo in range(0, N, width)]
    ref.steps(batches)
    e.finito_steps(*csr(batches))
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    assert rel(e.get_table_rows(), ref.s) < 1e-9
    z1 = e.get_vec(L.VEC_Z)
    e.finito_init(x0, gam, ref.hat_gamma)
    e.finito_steps(*csr(batches))
    assert np.array_equal(z1, e.get_vec(L.VEC_Z))
    e.close()


@pytest.mark.parametrize("exchange", ["words", "words_two_dots", "barrier"])
def test_lfinito_minibatch_many_small_batches(exchange, monkeypatch):
    """Many batches per launch (65 per sweep, 5 sweeps) with more CTAs than row groups: CTAs without rows still take part in
    the exchange of every batch; the epoch counter of the flagged words keeps counting across the launches.  By default the sweep
    takes c_i(z_full) of every row from the pass at z_full that opened it; "two_dots" forms it again from a second dot product."""
    monkeypatch.setenv("CIAO_BATCH_EXCHANGE", exchange.split("_")[0])
    if exchange.endswith("two_dots"):
        monkeypatch.setenv("CIAO_BATCH_TWO_DOTS", "1")
    kind, N, d, batch = orc.LOSS_LS, 16600, 130, 256
    p, e = make_rows(kind, N, d, 0xBAB, lam_reg=0.05)
    gam = 0.999 * N / (np.sum(p.A * p.A, axis=1) * N)
    x0 = np.full(d, 0.1)
    ref = orc.LFinitoState(p, x0, gam, batch)
    e.lfinito_init(x0, gam, ref.hat_gamma)
    sw = LFinitoSweeper(N, batch, 3, HostRNG(5))
    for _ in range(5):
        order = sw.next()
        ref.outer(order)
        e.lfinito_outer(order, batch)
    assert rel(e.get_vec(L.VEC_Z), ref.z) < 1e-9
    assert rel(e.get_vec(L.VEC_Z_FULL), ref.z_full) < 1e-9
    assert rel(e.get_vec(L.VEC_AV), ref.av) < 1e-8
    e.close()

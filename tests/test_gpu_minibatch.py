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
                                             (orc.LOSS_LS, 1024, 4096, 256), (orc.LOSS_LOGISTIC, 900, 130, 300)])
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


@pytest.mark.parametrize("kind,N,d,batch", [(orc.LOSS_LS, 1500, 64, 256), (orc.LOSS_LOGISTIC, 2100, 1024, 700)])
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

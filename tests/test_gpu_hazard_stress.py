"""Stress of the table read-after-write paths (SAGA_basic.jl:65, Finito_basic.jl:116, ProShI_basic.jl:119).

Table rows are written by the compute threads (st.global) and prefetched D steps ahead into a shared-memory ring by cp.async
copies of the producer warp (both generic proxy).  A repeat closer than CIAO_HAZARD_WINDOW = 20 steps carries a HAZARD flag and
is re-read by the thread that wrote it; a repeat further away is ordered by an mbarrier release/acquire chain at CTA scope
(seq_impl.cuh, proshi.cu).  These tests put
≥ 10^6 steps on tiny problems (N = 24, 40, 64), with every repeat distance around the window edges — 19, 20, 21 (window − 1,
window, window + 1), 9, 10 (SAGA/Finito ring + 1, + 2), 17, 18 (ProShI ring) — and require
  * SAGA / Finito: the TMA-ring path BITWISE equal to the register-prefetch path (CIAO_SEQ_TABLE_LDG=1), which never touches
    the async proxy;
  * ProShI: BITWISE equal to the CPU oracle (the block update is elementwise, so there is no summation-order licence).
A stale row anywhere in the 10^6 steps changes the bits of everything after it.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine

gpu = pytest.mark.gpu


def stress_indices(N, K, seed):
    """1-based index stream of length K over 1..N: cycles of period P over P distinct rows (every repeat at distance exactly
    P) for the edge periods and for random P in [9, min(N, 60)], interleaved with uniformly random stretches."""
    rng = np.random.default_rng(seed)
    edge = [p for p in (8, 9, 10, 11, 16, 17, 18, 19, 20, 21, 22) if p <= N]
    out, n = [], 0
    while n < K:
        kind = rng.integers(0, 4)
        if kind == 0:
            seg = rng.integers(1, N + 1, size=int(rng.integers(50, 400)))
        else:
            P = int(rng.choice(edge)) if kind < 3 else int(rng.integers(9, min(N, 60) + 1))
            rows = rng.permutation(N)[:P] + 1
            seg = np.tile(rows, int(rng.integers(3, 40)))
        out.append(seg.astype(np.int64))
        n += len(seg)
    return np.concatenate(out)[:K]


def distances(idx):
    last, dist = {}, np.zeros(len(idx), dtype=np.int64)
    for k, v in enumerate(idx.tolist()):
        dist[k] = k - last[v] if v in last else 0
        last[v] = k
    return dist


def test_stress_stream_covers_the_window_edges():
    idx = stress_indices(40, 200_000, 1)
    h = np.bincount(distances(idx), minlength=64)
    assert h[1] > 300
    for dd in (9, 10, 17, 18, 19, 20, 21, 22, 40):
        assert h[dd] > 1000, (dd, h[dd])


@gpu
@pytest.mark.parametrize("N", [24, 40, 64])
@pytest.mark.parametrize("d", [1024, 4096])
@pytest.mark.parametrize("alg", ["saga", "finito"])
def test_table_ring_bitwise_equals_register_path(N, d, alg, monkeypatch):
    K = 1_000_000
    idx = stress_indices(N, K, 100 * N + d)
    outs = []
    for ldg in ("0", "1"):
        monkeypatch.setenv("CIAO_SEQ_TABLE_LDG", ldg)    # read by ciao_create
        with Engine(0) as e:
            e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x57E55 + d, scale=float(N))
            e.set_reg(L.REG_NORML1, 0.05)
            Lmax = N * e.max_row_sqnorm()
            x0 = np.full(d, 0.125)
            if alg == "saga":
                e.saga_init(x0, 1 / (3 * Lmax), False)
                e.saga_steps(idx[: K // 2])
                e.saga_steps(idx[K // 2:])
            else:
                gam = np.linspace(0.7, 1.0, N) * (0.999 * N / Lmax)
                e.finito_init(x0, gam, 1 / np.sum(1 / gam))
                e.finito_steps(idx, np.arange(K + 1, dtype=np.int64))
            outs.append((e.get_vec(L.VEC_Z), e.get_vec(L.VEC_AV), e.get_table_rows()))
            assert np.all(np.isfinite(outs[-1][0]))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@gpu
@pytest.mark.parametrize("N", [24, 40, 64])
@pytest.mark.parametrize("n", [96, 1024])
def test_proshi_long_run_bitwise_equals_oracle(N, n):
    K = 1_000_000 if n == 96 else 200_000
    idx = stress_indices(N, K, 7 * N + n)
    ptr = np.arange(K + 1, dtype=np.int64)
    ptr = ptr[np.sort(np.unique(np.concatenate([[0, K], np.random.default_rng(3).integers(0, K, size=K // 2)])))]  # batches of 1-8
    Q, _ = orc.gen_rows(orc.SYN_SHARING, n, 0x5EED0005, 0, N)
    eta = 10.0 * N
    p = orc.Problem(orc.LOSS_DIAGQUAD, Q, np.ones((N, n)), box=(-2.0, 2.0), eta=eta).set_reg(orc.REG_INDBOX, lo=-np.inf, hi=np.ones(n))
    gam = 0.999 * N / (np.abs(Q).max(axis=1) + eta)
    ref = orc.ProshiState(p, np.zeros(n), gam)
    with Engine(0) as e:
        e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005)
        e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
        e.proshi_init(np.zeros(n), gam, ref.hat_gamma)
        # the init sums Σ s_i in different orders (Julia: left to right; device: per CTA, then slices) — a 1-ulp difference in
        # av would never wash out, so the run starts from the oracle's av; everything after is elementwise and must be exact
        av0 = e.get_vec(L.VEC_AV)
        assert np.abs(av0 - ref.av).max() <= 4e-16 * np.abs(ref.av).max() and np.array_equal(e.get_vec(L.VEC_Z), ref.z)
        e.set_vec(L.VEC_AV, ref.av)
        ref.steps(idx, ptr)
        e.proshi_steps(idx, ptr)
        assert np.array_equal(e.get_table_rows(), ref.s)
        assert np.array_equal(e.get_vec(L.VEC_Z), ref.z)
        assert np.array_equal(e.get_vec(L.VEC_AV), ref.av)

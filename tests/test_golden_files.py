"""The committed fixtures of tests/golden/: the reference's constants match tests/fixtures.py value for value, and the
oracle reproduces its committed trajectories (CPU); the engine reaches the same vectors (GPU)."""
import json
import os
import sys

import numpy as np
import pytest

import fixtures

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)


def test_reference_constants_match_fixtures():
    c = json.load(open(os.path.join(GOLD, "reference_constants.json")))
    lg, sh = fixtures.logistic_l1(), fixtures.sharing()
    assert np.array_equal(np.array(c["logistic_l1"]["A"]), lg["A"]) and np.array_equal(np.array(c["logistic_l1"]["x_star"]), lg["x_star"])
    assert np.array_equal(np.array(c["sharing"]["Qdiag"]), sh["Qdiag"]) and np.array_equal(np.array(c["sharing"]["sum_star"]), sh["sum_star"])
    assert c["logistic_l1"]["x_star_source"] == "test/test_logistic_l1.jl:29" and c["sharing"]["sum_star_source"] == "test/test_sharing.jl:28"


def test_oracle_reproduces_committed_trajectories():
    import make_golden
    gold = np.load(os.path.join(GOLD, "oracle_trajectories.npz"))
    now = make_golden.trajectories()
    assert sorted(gold.files) == sorted(now)
    for k in gold.files:   # bit-equal on the same CPU; libm's exp may differ by an ulp across CPU models (ifunc variants)
        a, b = np.asarray(gold[k], dtype=float), np.asarray(now[k], dtype=float)
        assert a.shape == b.shape and np.linalg.norm(a - b) <= 1e-13 * max(np.linalg.norm(a), 1e-300), k


@pytest.mark.gpu
def test_engine_reaches_committed_trajectories():
    from ciaoalgorithms_jl_b200 import _lib as L
    from ciaoalgorithms_jl_b200.engine import Engine
    gold = np.load(os.path.join(GOLD, "oracle_trajectories.npz"))
    fx = fixtures.logistic_l1()
    N, Lc, x0 = fx["N"], fx["L"], fx["x0"]
    gam = 0.999 * N / Lc

    def close(a, b):
        return np.linalg.norm(a - b) <= 1e-10 * max(np.linalg.norm(b), 1e-300)

    with Engine(0) as e:
        e.set_rows(L.LOSS_LOGISTIC, fx["A"], fx["y"], fx["mu"]); e.set_reg(L.REG_NORML1, fx["lam"])
        e.finito_init(x0, gam, 1 / np.sum(1 / gam))
        idx = gold["logistic_finito_cyclic_idx"]
        e.finito_steps(idx, np.arange(len(idx) + 1, dtype=np.int64))
        assert close(e.get_vec(L.VEC_Z), gold["logistic_finito_cyclic_z"])
        e.saga_init(x0, 1 / (3 * Lc.max()), False); e.saga_steps(gold["logistic_saga_idx"])
        assert close(e.get_vec(L.VEC_Z), gold["logistic_saga_z"])
        e.svrg_init(x0, 1 / (10 * Lc.max()), False)
        for ep in gold["logistic_svrg_idx"]:
            e.svrg_epoch(np.ascontiguousarray(ep))
        assert close(e.get_vec(L.VEC_Z_FULL), gold["logistic_svrg_z_full"])
    fl = fixtures.planted_lasso(0)
    with Engine(0) as e:
        e.set_rows(L.LOSS_LS, fl["A"], fl["b"], fl["scale"]); e.set_reg(L.REG_NORML1, fl["lam"])
        e.finito_adaptive_init(fl["x0"])
        assert e.finito_adaptive_steps(gold["lasso_adaptive_idx"]) == len(gold["lasso_adaptive_idx"])
        g, _, _, _, nbt = e.finito_adaptive_get()
        assert nbt == int(gold["lasso_adaptive_backtracks"][0]) and close(g, gold["lasso_adaptive_gamma"])
        assert close(e.get_vec(L.VEC_Z), gold["lasso_adaptive_z"])
    fs = fixtures.sharing()
    with Engine(0) as e:
        e.set_blocks(fs["Qdiag"], fs["qlin"], fs["box"], fs["eta"]); e.set_reg(L.REG_INDBOX, -np.inf, fs["g_hi"])
        gm = 0.999 * fs["N"] / fs["L"]
        e.proshi_init(fs["x0"], gm, float(np.sum(gm)))
        idx = gold["sharing_proshi_idx"]
        e.proshi_steps(idx, np.arange(len(idx) + 1, dtype=np.int64))
        assert close(e.get_vec(L.VEC_Z), gold["sharing_proshi_z"]) and close(e.get_table_rows(), gold["sharing_proshi_s"])

"""Writes the committed golden fixtures of tests/golden/.

  reference_constants.json  — the golden vectors and fixture data the REFERENCE's own tests hold for the hot path, copied
                              value for value with their file:line (the reference is Julia and cannot be executed in this
                              image or on the GPU box, so these are its only reference-produced numbers).
  oracle_trajectories.npz   — short trajectories of every solver variant on those fixtures, produced by the CPU oracle
                              (oracle/ciao_oracle.c) with fixed index sequences.  They lock the oracle against silent
                              changes (tests/test_golden_files.py, CPU) and give the GPU suite vectors to compare with that
                              do not depend on the oracle being built on the box.

    python tests/golden/make_golden.py        (from the repo root; rewrites both files)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ciao_pkg  # noqa: E402

ciao_pkg.load()
import fixtures  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from ciaoalgorithms_jl_b200.sampling import AdaptiveSweeper, BatchSweeper, HostRNG, LFinitoSweeper  # noqa: E402


def constants():
    lg, sh = fixtures.logistic_l1(), fixtures.sharing()
    return {
        "logistic_l1": {"source": "test/test_logistic_l1.jl:12-29", "A": lg["A"].tolist(), "y": lg["y"].tolist(),
                        "lambda": lg["lam"], "x0": lg["x0"].tolist(), "x_star": lg["x_star"].tolist(),
                        "x_star_source": "test/test_logistic_l1.jl:29", "tolerance": 1e-4},
        "sharing": {"source": "test/test_sharing.jl:9-28", "Qdiag": sh["Qdiag"].tolist(), "q": sh["qlin"].tolist(),
                    "box": list(sh["box"]), "eta": sh["eta"], "L": sh["L"].tolist(), "g_upper": sh["g_hi"].tolist(),
                    "sum_star": sh["sum_star"].tolist(), "sum_star_source": "test/test_sharing.jl:28", "tolerance": 1e-4},
        "lasso": {"source": "test/test_lasso.jl:15-60 (optimum planted by construction; instance rebuilt with numpy seed 0)",
                  "criterion": "cost(x) - f_star < 1e-4 after maxit = 1000 (SVRG++: 16 with m = 1; SAG: 10000)"},
    }


def trajectories():
    out = {}
    fx = fixtures.logistic_l1()
    p = orc.Problem(orc.LOSS_LOGISTIC, fx["A"], fx["y"], fx["mu"]).set_reg(orc.REG_NORML1, lam=fx["lam"])
    N, L, x0 = fx["N"], fx["L"], fx["x0"]
    gam = 0.999 * N / L
    st = orc.FinitoState(p, x0, gam); b = BatchSweeper(N, 1, 2, HostRNG(1)).take(40); st.steps(b)
    out["logistic_finito_cyclic_idx"] = np.concatenate(b); out["logistic_finito_cyclic_z"] = st.z.copy()
    st = orc.LFinitoState(p, x0, gam, 1); sw = LFinitoSweeper(N, 1, 2, HostRNG(1))
    for _ in range(5):
        st.outer(sw.next())
    out["logistic_lfinito_5_outer_z"] = st.z.copy()
    idx = HostRNG(3).rand_vec(N, 40)
    st = orc.SAGAState(p, x0, 1 / (3 * L.max())); st.steps(idx)
    out["logistic_saga_idx"] = idx; out["logistic_saga_z"] = st.z.copy()
    st = orc.SVRGState(p, x0, 1 / (10 * L.max()), m=N); ep = [HostRNG(5 + k).rand_vec(N, N) for k in range(3)]
    for e in ep:
        st.epoch(e)
    out["logistic_svrg_idx"] = np.stack(ep); out["logistic_svrg_z_full"] = st.z_full.copy()
    fl = fixtures.planted_lasso(0)
    pl = orc.Problem(orc.LOSS_LS, fl["A"], fl["b"], fl["scale"]).set_reg(orc.REG_NORML1, lam=fl["lam"])
    ia = AdaptiveSweeper(fl["N"], 2, HostRNG(1)).take(60)
    st = orc.FinitoAdaptiveState(pl, fl["x0"]); st.steps(ia)
    out["lasso_adaptive_idx"] = ia; out["lasso_adaptive_z"] = st.z.copy(); out["lasso_adaptive_gamma"] = st.gamma.copy()
    out["lasso_adaptive_backtracks"] = np.array([st.backtracks])
    fs = fixtures.sharing()
    ps = orc.Problem(orc.LOSS_DIAGQUAD, fs["Qdiag"], fs["qlin"], box=fs["box"], eta=fs["eta"]).set_reg(orc.REG_INDBOX, lo=-np.inf, hi=fs["g_hi"])
    st = orc.ProshiState(ps, fs["x0"], 0.999 * fs["N"] / fs["L"]); b = BatchSweeper(fs["N"], 1, 2, HostRNG(1)).take(30); st.steps(b)
    out["sharing_proshi_idx"] = np.concatenate(b); out["sharing_proshi_z"] = st.z.copy(); out["sharing_proshi_s"] = st.s.copy()
    return out


if __name__ == "__main__":
    json.dump(constants(), open(os.path.join(HERE, "reference_constants.json"), "w"), indent=1)
    np.savez(os.path.join(HERE, "oracle_trajectories.npz"), **trajectories())
    print("wrote", os.listdir(HERE))

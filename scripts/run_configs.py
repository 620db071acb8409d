"""Measures BASELINE.json configs[1] (C2: L1-logistic N=2^20 d=1024, SAGA + Finito) and configs[4]
(C5: sharing N=2^18 blocks n=1024, ProShI) on one B200, event-timed, with an oracle-side CPU sample.
Writes gpurun_out/configs_r2.json (copied to profiles/ by hand)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ciao_pkg  # noqa: E402

ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L  # noqa: E402
from ciaoalgorithms_jl_b200.engine import Engine  # noqa: E402
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, csr  # noqa: E402
from oracle import oracle as orc  # noqa: E402

small = "--small" in sys.argv
out = {}

# ------------------------------------------------------------------ C2: logistic-L1, SAGA + Finito
N, d = (1 << 16, 1024) if small else (1 << 20, 1024)
e = Engine(0)
e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0)
e.set_reg(L.REG_NORML1, 1.0 / N)
Lmax = 0.25 * e.max_row_sqnorm()
x0 = np.ones(d)
ld = d + 8
rng = HostRNG(0x1D0002)
c2 = {"N": N, "d": d}
# SAGA
e.saga_init(x0, 1 / (3 * Lmax), False)
t = e.last_timing()
c2["saga_table_init"] = {"ms": t.last_pass_ms, "GBs": t.last_pass_bytes / t.last_pass_ms / 1e6, "bytes": t.last_pass_bytes}
f0 = sum(e.objective(x0))
tot_ms, K = 0.0, 3
for ep in range(K):
    idx = rng.rand_vec(N, N)
    e.saga_steps(idx)
    tot_ms += e.last_timing().last_seq_ms
z = e.get_vec(L.VEC_Z)
c2["saga"] = {"epochs": K, "us_per_step": 1e3 * tot_ms / (K * N), "epochs_per_s": K / (tot_ms / 1e3), "objective0": f0,
              "objective": sum(e.objective(z)), "bytes_per_step": 24 * d + 32}
# Finito: sweeping 1 (random) and 2 (cyclic), batch 1
Li = None
for sweeping in (1, 2):
    # L_i = 0.25‖a_i‖² needs the row norms: use the global bound for all i (scalar L, Finito_basic.jl:68)
    gam = np.full(N, 0.999 * N / Lmax)
    hat = 1 / np.sum(1 / gam)
    e.finito_init(x0, gam, hat)
    t = e.last_timing()
    c2.setdefault("finito_table_init", {"ms": t.last_pass_ms, "GBs": t.last_pass_bytes / t.last_pass_ms / 1e6})
    tot_ms = 0.0
    sw = BatchSweeper(N, 1, sweeping, rng)
    for ep in range(K):
        idx = rng.rand_vec(N, N) if sweeping == 1 else np.concatenate([np.arange(2, N + 1), [1]]).astype(np.int64)
        bp = np.arange(N + 1, dtype=np.int64)
        e.finito_steps(idx, bp)
        tot_ms += e.last_timing().last_seq_ms
    c2[f"finito_sweeping{sweeping}"] = {"epochs": K, "us_per_step": 1e3 * tot_ms / (K * N), "epochs_per_s": K / (tot_ms / 1e3),
                                        "objective": sum(e.objective(e.get_vec(L.VEC_Z)))}
# Finito / LFinito with static minibatches of 4096 rows: streaming passes (batch.cu)
gam = np.full(N, 0.999 * N / Lmax)
hat = 1 / np.sum(1 / gam)
e.finito_init(x0, gam, hat)
sw = BatchSweeper(N, 4096, 2, rng)
tot_ms = 0.0
for ep in range(K):
    idx, bp = csr(sw.take(sw.d))
    e.finito_steps(idx, bp)
    tot_ms += e.last_timing().last_seq_ms
c2["finito_cyclic_batch4096"] = {"epochs": K, "us_per_row": 1e3 * tot_ms / (K * N), "epochs_per_s": K / (tot_ms / 1e3),
                                 "GBs": K * N * (ld + 2 * d) * 8 / tot_ms / 1e6, "objective": sum(e.objective(e.get_vec(L.VEC_Z)))}
e.lfinito_init(x0, gam, hat)
tot_ms = 0.0
for ep in range(K):
    e.lfinito_outer(np.arange(1, sw.d + 1), 4096)
    tot_ms += e.last_timing().last_seq_ms
c2["lfinito_sweep_batch4096"] = {"sweeps": K, "us_per_row": 1e3 * tot_ms / (K * N), "sweeps_per_s": K / (tot_ms / 1e3),
                                 "GBs": K * N * ld * 8 / tot_ms / 1e6, "objective": sum(e.objective(e.get_vec(L.VEC_Z)))}
e.close()
# CPU sample (oracle, 1 thread): SAGA steps on 2^14 rows
Ns = 1 << 14
A, y = orc.gen_rows(orc.SYN_LOGISTIC, d, 0x5EED0002, 0, Ns)
p = orc.Problem(orc.LOSS_LOGISTIC, A, y, np.ones(Ns)).set_reg(orc.REG_NORML1, lam=1.0 / Ns)
st = orc.SAGAState(p, np.ones(d), 1 / (3 * 0.25 * p.max_row_sqnorm()))
idx = HostRNG(1).rand_vec(Ns, 4 * Ns)
t0 = time.perf_counter()
st.steps(idx)
c2["cpu_saga_us_per_step"] = 1e6 * (time.perf_counter() - t0) / len(idx)
out["C2_logistic_saga_finito"] = c2
print(json.dumps(c2), flush=True)

# ------------------------------------------------------------------ C5: sharing, ProShI
N, n = (1 << 14, 1024) if small else (1 << 18, 1024)
e = Engine(0)
e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005)
e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
eta = 10.0 * N
Li = np.full(N, 10.0 + eta)            # max_j |q_ij| + η ≤ 10 + η  (scalar bound; the generator draws q in (-1, 10))
gam = 0.999 * N / Li
c5 = {"N": N, "n": n}
e.proshi_init(np.zeros(n), gam, float(np.sum(gam)))
t = e.last_timing()
c5["table_init"] = {"ms": t.last_pass_ms, "GBs": t.last_pass_bytes / t.last_pass_ms / 1e6}
rng = HostRNG(0x1D0005)
for r in (1, 4096):
    sw = BatchSweeper(N, r, 2, rng)
    tot_ms, K = 0.0, 3
    for ep in range(K):
        batches = sw.take(sw.d)
        idx, bp = csr(batches)
        e.proshi_steps(idx, bp)
        tot_ms += e.last_timing().last_seq_ms
    # GBs: SURVEY §8d's algorithmic figure (24n per block: table read + write, diag(Q_i)); dram_GBs: what the general block layout
    # moves (q_i, c_i, s_i read + s_i written = 32n; ncu: 6.47 GB read + 2.10 GB written per sweep)
    c5[f"proshi_cyclic_batch{r}"] = {"sweeps": K, "us_per_block": 1e3 * tot_ms / (K * N), "GBs": 24.0 * n * N * K / tot_ms / 1e6,
                                     "dram_GBs": 32.0 * n * N * K / tot_ms / 1e6, "sweeps_per_s": K / (tot_ms / 1e3)}
sw = BatchSweeper(N, 1, 1, rng)
idx = rng.rand_vec(N, N)
e.proshi_steps(idx, np.arange(N + 1, dtype=np.int64))
ms = e.last_timing().last_seq_ms
c5["proshi_random_batch1"] = {"us_per_block": 1e3 * ms / N, "GBs": 24.0 * n * N / ms / 1e6}
e.proshi_solution(None)
t = e.last_timing()
c5["solution_inplace"] = {"ms": t.last_pass_ms, "GBs": t.last_pass_bytes / t.last_pass_ms / 1e6}
c5["sum_x_first3"] = e.table_colsum()[:3].tolist()
e.close()
Ns = 1 << 12
Q, _ = orc.gen_rows(orc.SYN_SHARING, n, 0x5EED0005, 0, Ns)
p = orc.Problem(orc.LOSS_DIAGQUAD, Q, np.ones((Ns, n)), box=(-2.0, 2.0), eta=10.0 * Ns).set_reg(orc.REG_INDBOX, lo=-np.inf, hi=np.ones(n))
gm = 0.999 * Ns / np.full(Ns, 10.0 + 10.0 * Ns)
st = orc.ProshiState(p, np.zeros(n), gm)
batches = BatchSweeper(Ns, 1, 2, HostRNG(1)).take(4 * Ns)
t0 = time.perf_counter()
st.steps(batches)
c5["cpu_proshi_us_per_block"] = 1e6 * (time.perf_counter() - t0) / len(batches)
out["C5_sharing_proshi"] = c5
print(json.dumps(c5), flush=True)

# ------------------------------------------------------------------ adaptive Finito (Finito_adaptive.jl) on a Lasso, N = 2^18, d = 1024
from ciaoalgorithms_jl_b200.sampling import AdaptiveSweeper  # noqa: E402
N, d = (1 << 14, 1024) if small else (1 << 18, 1024)
e = Engine(0)
e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N))
e.set_reg(L.REG_NORML1, N / 100.0)
x0 = np.full(d, 0.01)
t0 = time.perf_counter()
e.finito_adaptive_init(x0, 0.999, 1e-9)
init_s = time.perf_counter() - t0
f0 = sum(e.objective(x0))
idx = AdaptiveSweeper(N, 1, HostRNG(0x1D0006)).take(2 * N)
done = e.finito_adaptive_steps(idx)
ms = e.last_timing().last_seq_ms
_, _, _, hat, nbt = e.finito_adaptive_get(gamma=False)
ad = {"N": N, "d": d, "init_s": init_s, "steps": done, "us_per_step": 1e3 * ms / max(done, 1), "linesearch_reductions": nbt,
      "objective0": f0, "objective": sum(e.objective(e.get_vec(L.VEC_Z))), "hat_gamma": hat}
e.close()
out["adaptive_finito_lasso"] = ad
print(json.dumps(ad), flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_r2.json"), "w"), indent=1)

"""Sweeps the pass tuning knobs on the C2 table-init passes (SAGA / Finito init, N = 2^20, d = 1024)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine
d = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = (1 << 30) // d
e = Engine(0); e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0); e.set_reg(L.REG_NORML1, 1.0 / N)
Lmax = 0.25 * e.max_row_sqnorm()
gam = np.full(N, 0.999 * N / Lmax); hat = 1 / np.sum(1 / gam)
combos = [(0, 0, 0)] + [(T, S, C) for T in (64, 128, 256) for S in (2, 3, 4) for C in (2, 3, 4, 5, 6, 8)]
for T, S, C in combos:
    try:
        e.set_tuning(pass_threads=T, pass_stages=S, pass_ctas_per_sm=C)
        ts, tf, tg = [], [], []
        for _ in range(3):
            e.saga_init(np.ones(d), 1 / (3 * Lmax), False); ts.append(e.last_timing().last_pass_ms)
            e.finito_init(np.ones(d), gam, hat); tf.append(e.last_timing().last_pass_ms)
            e.full_gradient(np.ones(d), 1.0, out=False); tg.append(e.last_timing().last_pass_ms)
        print(f"T={T} S={S} C={C}: saga init {min(ts):.3f} ms, finito init {min(tf):.3f} ms, full gradient {min(tg):.3f} ms ({N * (d + 8) * 8 / min(tg) / 1e6:.0f} GB/s)", flush=True)
    except Exception as ex:
        pass

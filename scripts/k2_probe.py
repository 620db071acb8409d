import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine
N, d = 1 << 20, 1024
e = Engine(0); e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0); e.set_reg(L.REG_NORML1, 1.0 / N)
Lmax = 0.25 * e.max_row_sqnorm()
for _ in range(3):
    e.saga_init(np.ones(d), 1 / (3 * Lmax), False)
    t = e.last_timing(); print("saga init", t.last_pass_ms, t.last_pass_bytes / t.last_pass_ms / 1e6)
gam = np.full(N, 0.999 * N / Lmax)
for _ in range(3):
    e.finito_init(np.ones(d), gam, 1 / np.sum(1 / gam))
    t = e.last_timing(); print("finito init", t.last_pass_ms, t.last_pass_bytes / t.last_pass_ms / 1e6)

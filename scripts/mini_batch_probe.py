import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, csr
N, d = 1 << 16, 1024
e = Engine(0)
e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0); e.set_reg(L.REG_NORML1, 1.0 / N)
Lmax = 0.25 * e.max_row_sqnorm()
gam = np.full(N, 0.999 * N / Lmax); hat = 1 / np.sum(1 / gam)
e.finito_init(np.ones(d), gam, hat)
sw = BatchSweeper(N, 4096, 2, HostRNG(1))
for _ in range(2):
    idx, bp = csr(sw.take(sw.d)); e.finito_steps(idx, bp)
e.sync(); print("finito ms per batch", e.last_timing().last_seq_ms / sw.d)
e.lfinito_init(np.ones(d), gam, hat)
for _ in range(2):
    e.lfinito_outer(np.arange(1, sw.d + 1), 4096)
e.sync(); print("lfinito ms per batch", e.last_timing().last_seq_ms / sw.d)

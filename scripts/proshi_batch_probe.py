import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
if os.environ.get("CIAO_SO"): L.SO_PATH = os.environ["CIAO_SO"]
from ciaoalgorithms_jl_b200.engine import Engine
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, csr
N, n = 1 << 18, 1024
e = Engine(0); e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005); e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
gam = 0.999 * N / np.full(N, 10.0 + 10.0 * N)
for r in (256, 1024, 4096, 16384):
    e.proshi_init(np.zeros(n), gam, float(np.sum(gam)))
    sw = BatchSweeper(N, r, 2, HostRNG(1)); idx, bp = csr(sw.take(sw.d))
    e.proshi_steps(idx, bp); e.proshi_steps(idx, bp)
    ms = e.last_timing().last_seq_ms
    print(os.environ.get("CIAO_SO", "default")[-12:], "batch", r, f"{1e3*ms/N:.5f} us/block  {24.0*n*N/ms/1e6:.0f} GB/s")

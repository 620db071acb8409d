import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
if os.environ.get("CIAO_SO"): L.SO_PATH = os.environ["CIAO_SO"]
from ciaoalgorithms_jl_b200.engine import Engine
N, n = 1 << 18, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
e = Engine(0); e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005); e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
gam = 0.999 * N / np.full(N, 10.0 + 10.0 * N)
e.proshi_init(np.zeros(n), gam, float(np.sum(gam)))
idx = np.random.default_rng(1).integers(1, N + 1, size=N, dtype=np.int64); bp = np.arange(N + 1, dtype=np.int64)
e.proshi_steps(idx, bp); e.proshi_steps(idx, bp)
print(os.environ.get("CIAO_SO", "default")[-10:], f"n={n} batch 1: {1e3 * e.last_timing().last_seq_ms / N:.4f} us/block")

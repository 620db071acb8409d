"""Per-call overhead of the step ABI with K = 1 (the `iterator` protocol: one ccall per iterate)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine
N, d = 1 << 16, 1024
e = Engine(0); e.gen_synthetic(L.SYNTH_LASSO, N, d, 3, scale=float(N)); e.set_reg(L.REG_NORML1, N / 100.0)
g = 1.0 / (7.0 * N * e.max_row_sqnorm())
idx = np.random.default_rng(0).integers(1, N + 1, size=4096, dtype=np.int64)
e.saga_init(np.zeros(d), g, False)
for K in (1, 16, 256, 4096):
    e.saga_steps(idx[:K]); e.sync()
    t0 = time.perf_counter()
    reps = 200 if K < 4096 else 20
    for r in range(reps):
        e.saga_steps(idx[:K])
    e.sync()
    dt = (time.perf_counter() - t0) / reps
    print(f"saga_steps K={K}: {1e6 * dt:.1f} us per call, {1e6 * dt / K:.2f} us per step")
e.svrg_init(np.zeros(d), g, False)
t0 = time.perf_counter()
for r in range(20):
    e.svrg_epoch(idx[:256])
e.sync()
print(f"svrg_epoch m=256 (N={N}): {1e6 * (time.perf_counter() - t0) / 20:.1f} us per call (includes a full-gradient pass over {N} rows)")

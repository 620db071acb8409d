"""Per-phase cycle profile of the sequential kernel (needs the CIAO_SEQ_PROFILE build: libciao_cuda_prof.so)."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
L.SO_PATH = os.path.join(ROOT, "ciaoalgorithms.jl_b200", "libciao_cuda_prof.so")
from ciaoalgorithms_jl_b200.engine import Engine
import torch
rows_log2, d = int(sys.argv[1]), int(sys.argv[2])
N = 1 << rows_log2
e = Engine(0)
e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N)); e.set_reg(L.REG_NORML1, N / 100.0)
gamma = 1.0 / (7.0 * N * e.max_row_sqnorm())
m = min(N, 1 << 17)
idx = np.random.default_rng(1).integers(1, N + 1, size=m, dtype=np.int64)
names = ["wait_row", "lds_dot_shfl", "exchange", "sum_update"]
prof = (C.c_longlong * 4)()
for C_, T in [(8, 128), (4, 256), (8, 64), (8, 256), (4, 128)]:
    e.set_tuning(seq_cluster=C_, seq_threads=T)
    e.svrg_init(np.zeros(d), gamma, True)
    e.svrg_epoch(idx)
    t = e.last_timing()
    e.lib.ciao_debug_seq_prof(e.h, prof)
    tot = sum(prof)
    print(f"C={C_} T={T}: {1e3 * t.last_seq_ms / m:.3f} us/step; cycles/step: " +
          ", ".join(f"{n} {v / m:.0f}" for n, v in zip(names, prof)) + f", total {tot / m:.0f}")

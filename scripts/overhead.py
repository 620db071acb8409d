"""Where does the time between kernels go?  Per-call event/wall timing of svrg_epoch with device vs host indices."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ciao_pkg; ciao_pkg.load()
import torch
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine
rows_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 22
N, d = 1 << rows_log2, 4096
e = Engine(0)
e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N)); e.set_reg(L.REG_NORML1, N / 100.0)
gamma = 1.0 / (7.0 * N * e.max_row_sqnorm())
rng = np.random.default_rng(1)
for mode in ("device", "host", "device", "host"):
    e.svrg_init(np.zeros(d), gamma, True)
    e.sync()
    for k in range(4):
        m = (N // 16) << k
        idx = rng.integers(1, N + 1, size=m, dtype=np.int64)
        if mode == "device":
            t = torch.from_numpy(idx).cuda(); torch.cuda.synchronize(); arg = t.data_ptr()
        else:
            arg = idx
        w0 = time.perf_counter(); e.timer_begin()
        e.svrg_epoch(arg, m)
        w1 = time.perf_counter()
        ms = e.timer_end(); w2 = time.perf_counter()
        tm = e.last_timing()
        print(f"{mode:6s} m={m:8d}: call returns after {1e3*(w1-w0):8.2f} ms, events {ms:9.2f} ms, wall {1e3*(w2-w0):9.2f} ms, "
              f"inner {tm.last_seq_ms:9.2f} + pass {tm.last_pass_ms:6.2f} = {tm.last_seq_ms+tm.last_pass_ms:9.2f}  -> gap {ms-tm.last_seq_ms-tm.last_pass_ms:7.2f} ms")

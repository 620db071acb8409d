import os, sys, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
L.SO_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ciaoalgorithms.jl_b200", "libciao_cuda_prof.so")
from ciaoalgorithms_jl_b200.engine import Engine
from ciaoalgorithms_jl_b200.sampling import AdaptiveSweeper, HostRNG
N, d = 1 << 18, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
e = Engine(0); e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N)); e.set_reg(L.REG_NORML1, N / 100.0)
e.finito_adaptive_init(np.full(d, 0.01), 0.999, 1e-9)
idx = AdaptiveSweeper(N, 1, HostRNG(6)).take(N)
done = e.finito_adaptive_steps(idx); ms = e.last_timing().last_seq_ms
prof = (C.c_longlong * 128)(); fn = e.lib.ciao_debug_seq_prof_adaptive; fn.argtypes = [C.c_void_p, C.c_void_p]; fn(e.h, prof)
names = ["setup", "partials+send", "shadow", "exchange_wait", "sums+model_test", "main_update", "whole_step"]
print(f"adaptive d={d}: {1e3 * ms / done:.4f} us/step; cycles/step: " + ", ".join(f"{n} {prof[i] / done:.0f}" for i, n in enumerate(names) if n != "-"))

"""Per-phase cycle profile of the sequential kernels (needs the CIAO_SEQ_PROFILE build: libciao_cuda_prof.so,
`python ciaoalgorithms.jl_b200/build.py --profile`).

    python scripts/prof_seq.py <log2 rows> <d> [svrg,saga,finito] [ls,logistic]
"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
L.SO_PATH = os.environ.get("CIAO_SO", os.path.join(ROOT, "ciaoalgorithms.jl_b200", "libciao_cuda_prof.so"))
from ciaoalgorithms_jl_b200.engine import Engine
rows_log2, d = int(sys.argv[1]), int(sys.argv[2])
algs = (sys.argv[3] if len(sys.argv) > 3 else "svrg").split(",")
losses = (sys.argv[4] if len(sys.argv) > 4 else "ls").split(",")
shapes = [(8, 128)] if len(sys.argv) > 5 else [(8, 128), (4, 256), (8, 64), (8, 256), (4, 128)]
N = 1 << rows_log2
names = ["dot_shfl", "send_prefetch", "exchange_wait", "sum_update", "(row_wait)", "(producer_busy)"]
prof = (C.c_longlong * 128)()
per_cta = "--per-cta" in sys.argv
for loss in losses:
    e = Engine(0)
    if loss == "ls":
        e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N)); e.set_reg(L.REG_NORML1, N / 100.0)
        Lmax = N * e.max_row_sqnorm()
    else:
        e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0); e.set_reg(L.REG_NORML1, 1.0 / N)
        Lmax = 0.25 * e.max_row_sqnorm()
    m = min(N, 1 << 17)
    idx = np.random.default_rng(1).integers(1, N + 1, size=m, dtype=np.int64)
    x0 = np.zeros(d) if loss == "ls" else np.ones(d)
    for alg in algs:
        for C_, T in shapes:
            e.set_tuning(seq_cluster=C_, seq_threads=T)
            if alg == "svrg":
                e.svrg_init(x0, 1.0 / (7.0 * Lmax), True)
                e.svrg_epoch(idx); e.svrg_epoch(idx)
            elif alg in ("saga", "sag"):
                e.saga_init(x0, 1.0 / ((16.0 if alg == "sag" else 3.0) * Lmax), alg == "sag")
                e.saga_steps(idx); e.saga_steps(idx)
            elif alg == "lfinito":
                gam = np.full(N, 0.999 * N / Lmax)
                e.lfinito_init(x0, gam, 1 / np.sum(1 / gam))
                e.lfinito_outer(np.arange(1, N + 1, dtype=np.int64), 1)
                e.lfinito_outer(np.arange(1, N + 1, dtype=np.int64), 1)
            else:
                gam = np.full(N, 0.999 * N / Lmax)
                e.finito_init(x0, gam, 1 / np.sum(1 / gam))
                bp = np.arange(m + 1, dtype=np.int64)
                e.finito_steps(idx, bp); e.finito_steps(idx, bp)
            t = e.last_timing()
            steps = N if alg == "lfinito" else m
            fn = getattr(e.lib, f"ciao_debug_seq_prof_{'saga' if alg == 'sag' else alg}", None)
            line = f"{alg:7s} {loss:8s} d={d} C={C_} T={T}: {1e3 * t.last_seq_ms / steps:.4f} us/step"
            if fn is not None:
                fn.argtypes = [C.c_void_p, C.c_void_p]
                fn(e.h, prof)
                line += "; cycles/step: " + ", ".join(f"{n} {v / m:.0f}" for n, v in zip(names, prof)) + f", total {sum(prof[:4]) / m:.0f}"
                if per_cta:
                    for r in range(C_ if C_ else 8):
                        line += f"\n      CTA {r}: " + ", ".join(f"{n} {prof[8 * r + j] / m:.0f}" for j, n in enumerate(names))
            print(line, flush=True)
    e.close()

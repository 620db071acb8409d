"""µs/step of the sequential kernels for experiment builds (scripts/build_variants.py) and launch shapes.
    python scripts/seq_variants.py                 # all variants, one subprocess each
    python scripts/seq_variants.py --worker <tag>  # internal
"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(tag):
    import ciao_pkg; ciao_pkg.load()
    from ciaoalgorithms_jl_b200 import _lib as L
    if tag != "base":
        L.SO_PATH = os.path.join(ROOT, "ciaoalgorithms.jl_b200", f"libciao_cuda_{tag}.so")
    from ciaoalgorithms_jl_b200.engine import Engine
    res = []
    # SVRG++ at d = 4096 (the headline's inner kernel)
    N, d = 1 << 20, 4096
    e = Engine(0)
    e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N)); e.set_reg(L.REG_NORML1, N / 100.0)
    Lmax = N * e.max_row_sqnorm()
    idx = np.random.default_rng(1).integers(1, N + 1, size=1 << 19, dtype=np.int64)
    shapes = [(0, 0)] + ([(8, 64), (4, 128), (4, 256), (16, 64)] if tag == "base" else [])
    for C_, T in shapes:
        e.set_tuning(seq_cluster=C_, seq_threads=T)
        e.svrg_init(np.zeros(d), 1.0 / (7.0 * Lmax), True)
        e.svrg_epoch(idx)
        ts = []
        for _ in range(3):
            e.svrg_epoch(idx); ts.append(e.last_timing().last_seq_ms)
        res.append(f"svrg d=4096 C={C_} T={T}: {1e3 * min(ts) / len(idx):.4f} us/step smid {e.last_seq_placement()}")
    e.close()
    # SAGA / Finito at d = 1024, LS and logistic (table kernels: the proxy fence + written barrier sit here)
    N, d = 1 << 20, 1024
    for loss in ("ls", "logistic"):
        e = Engine(0)
        if loss == "ls":
            e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N)); e.set_reg(L.REG_NORML1, N / 100.0)
            Lmax = N * e.max_row_sqnorm(); x0 = np.zeros(d)
        else:
            e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0); e.set_reg(L.REG_NORML1, 1.0 / N)
            Lmax = 0.25 * e.max_row_sqnorm(); x0 = np.ones(d)
        idx = np.random.default_rng(2).integers(1, N + 1, size=1 << 19, dtype=np.int64)
        e.saga_init(x0, 1 / (3 * Lmax), False)
        e.saga_steps(idx)
        ts = []
        for _ in range(3):
            e.saga_steps(idx); ts.append(e.last_timing().last_seq_ms)
        res.append(f"saga {loss} d=1024: {1e3 * min(ts) / len(idx):.4f} us/step")
        gam = np.full(N, 0.999 * N / Lmax)
        e.finito_init(x0, gam, 1 / np.sum(1 / gam))
        bp = np.arange(len(idx) + 1, dtype=np.int64)
        e.finito_steps(idx, bp)
        ts = []
        for _ in range(3):
            e.finito_steps(idx, bp); ts.append(e.last_timing().last_seq_ms)
        res.append(f"finito {loss} d=1024: {1e3 * min(ts) / len(idx):.4f} us/step")
        e.close()
    # ProShI batch 1 (C5 shape)
    N, n = 1 << 18, 1024
    e = Engine(0)
    e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005); e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
    gam = 0.999 * N / np.full(N, 10.0 + 10.0 * N)
    e.proshi_init(np.zeros(n), gam, float(np.sum(gam)))
    idx = np.random.default_rng(3).integers(1, N + 1, size=N, dtype=np.int64)
    bp = np.arange(N + 1, dtype=np.int64)
    e.proshi_steps(idx, bp)
    ts = []
    for _ in range(3):
        e.proshi_steps(idx, bp); ts.append(e.last_timing().last_seq_ms)
    res.append(f"proshi batch1 n=1024: {1e3 * min(ts) / N:.4f} us/block")
    e.proshi_solution(None); e.proshi_solution(None)
    t = e.last_timing()
    res.append(f"proshi solution: {t.last_pass_ms:.3f} ms {t.last_pass_bytes / t.last_pass_ms / 1e6:.0f} GB/s")
    e.close()
    for r in res:
        print(f"[{tag}] {r}", flush=True)


if __name__ == "__main__":
    if "--worker" in sys.argv:
        worker(sys.argv[sys.argv.index("--worker") + 1])
    else:
        tags = ["base"] + sorted(f[len("libciao_cuda_"):-3] for f in os.listdir(os.path.join(ROOT, "ciaoalgorithms.jl_b200"))
                                 if f.startswith("libciao_cuda_") and f.endswith(".so") and "prof" not in f)
        for t in tags:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", t], timeout=600)

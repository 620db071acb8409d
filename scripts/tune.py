"""Sweeps the tuning knobs of the streaming pass and the sequential kernel on one GPU (event-timed)."""
import os, sys, json, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
if os.environ.get("CIAO_SO"): L.SO_PATH = os.environ["CIAO_SO"]
from ciaoalgorithms_jl_b200.engine import Engine

rows_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 22
d = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
what = sys.argv[3] if len(sys.argv) > 3 else "pass,seq"
N = 1 << rows_log2
e = Engine(0)
e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N))
e.set_reg(L.REG_NORML1, N / 100.0)
ld = (d + 3) // 4 * 4 + 8
res = []
if "pass" in what:
    e.set_vec(L.VEC_X, np.full(d, 1e-3))
    for T, S, C in [(256, 0, 1), (128, 0, 1), (512, 0, 1), (256, 3, 2), (128, 3, 2), (128, 2, 3), (256, 4, 1), (256, 2, 1), (512, 3, 2), (64, 2, 3)]:
        try:
            e.set_tuning(pass_threads=T, pass_stages=S, pass_ctas_per_sm=C)
            for _ in range(2):
                e.full_gradient(None, 1.0, out=False)
            ts = []
            for _ in range(5):
                e.full_gradient(None, 1.0, out=False)
                ts.append(e.last_timing().last_pass_ms)
            ms = float(np.median(ts))
            r = {"kernel": "pass", "threads": T, "stages": S, "ctas": C, "ms": ms, "GBs": N * ld * 8 / ms / 1e6}
        except Exception as ex:
            r = {"kernel": "pass", "threads": T, "stages": S, "ctas": C, "error": str(ex)}
        print(json.dumps(r), flush=True)
        res.append(r)
if "seq" in what:
    e.set_tuning()
    gamma = 1.0 / (7.0 * N * e.max_row_sqnorm())
    m = min(N, 1 << 18)
    idx = np.random.default_rng(1).integers(1, N + 1, size=m, dtype=np.int64)
    for C, T in [(0, 0), (16, 64), (16, 128), (16, 256), (8, 64), (8, 128), (8, 256), (4, 128), (4, 256), (2, 256)]:
        try:
            e.set_tuning(seq_cluster=C, seq_threads=T)
            e.svrg_init(np.zeros(d), gamma, True)
            e.svrg_epoch(idx)
            e.svrg_epoch(idx)
            t = e.last_timing()
            r = {"kernel": "svrg", "cluster": C, "threads": T, "ms": t.last_seq_ms, "us_per_step": 1e3 * t.last_seq_ms / m}
        except Exception as ex:
            r = {"kernel": "svrg", "cluster": C, "threads": T, "error": str(ex)}
        print(json.dumps(r), flush=True)
        res.append(r)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"tune_{rows_log2}_{d}.json"), "w"), indent=1)

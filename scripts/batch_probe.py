"""C2 minibatch sweeps (Finito / LFinito, batch 4096 and 512) through the persistent kernel; shape knobs via env."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
if os.environ.get("CIAO_SO"):
    L.SO_PATH = os.environ["CIAO_SO"]
from ciaoalgorithms_jl_b200.engine import Engine
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, csr
N, d = 1 << 20, 1024
e = Engine(0); e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0); e.set_reg(L.REG_NORML1, 1.0 / N)
Lmax = 0.25 * e.max_row_sqnorm(); gam = np.full(N, 0.999 * N / Lmax); hat = 1 / np.sum(1 / gam)
tag = " ".join(f"{k[11:]}={os.environ[k]}" for k in ("CIAO_BATCH_T", "CIAO_BATCH_CTAS", "CIAO_BATCH_STAGES", "CIAO_BATCH_PER_LAUNCH") if k in os.environ) or "default"
for r in [int(v) for v in os.environ.get('CIAO_PROBE_BATCHES', '4096,512').split(',')]:
    e.finito_init(np.ones(d), gam, hat)
    sw = BatchSweeper(N, r, 2, HostRNG(1)); idx, bp = csr(sw.take(sw.d))
    e.finito_steps(idx, bp); e.finito_steps(idx, bp); tf = e.last_timing().last_seq_ms
    e.lfinito_init(np.ones(d), gam, hat)
    e.lfinito_outer(np.arange(1, sw.d + 1), r); e.lfinito_outer(np.arange(1, sw.d + 1), r); tl = e.last_timing().last_seq_ms
    print(f"[{tag}] batch {r}: finito {1e3 / tf:.1f} epochs/s ({1e3 * tf / sw.d:.1f} us/batch), lfinito {1e3 / tl:.1f} sweeps/s ({1e3 * tl / sw.d:.1f} us/batch)", flush=True)

fn = getattr(e.lib, "ciao_debug_batch_prof", None)
if fn is not None:       # profile build (CIAO_SO=…libciao_cuda_prof.so): cycles per batch of CTA 0 in each phase of the last call
    import ctypes as C
    names = ["rows", "partial_write", "barrier1", "reduction", "barrier2", "z_reload"]
    for mode in ("finito", "lfinito"):
        sw = BatchSweeper(N, 4096, 2, HostRNG(1)); idx, bp = csr(sw.take(sw.d))
        if mode == "finito":
            e.finito_init(np.ones(d), gam, hat); e.finito_steps(idx, bp)
        else:
            e.lfinito_init(np.ones(d), gam, hat); e.lfinito_outer(np.arange(1, sw.d + 1), 4096)
        out = (C.c_longlong * 8)()
        fn.argtypes = [C.c_void_p, C.c_void_p]; fn(e.h, out)
        print(f"{mode} batch 4096, cycles per batch (CTA 0): " + ", ".join(f"{n} {out[i] / sw.d:.0f}" for i, n in enumerate(names)), flush=True)

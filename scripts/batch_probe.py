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
tag = " ".join(f"{k[11:]}={os.environ[k]}" for k in ("CIAO_BATCH_T", "CIAO_BATCH_CTAS", "CIAO_BATCH_STAGES", "CIAO_BATCH_PER_LAUNCH", "CIAO_BATCH_EXCHANGE", "CIAO_BATCH_STAGE_TABLE", "CIAO_BATCH_GROUP", "CIAO_BATCH_XPF", "CIAO_BATCH_TWO_DOTS") if k in os.environ) or "default"
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
    names = ["rows", "partial_write", "barrier1", "reduction", "barrier2", "z_reload"] if os.environ.get("CIAO_BATCH_EXCHANGE") == "barrier" \
        else ["rows", "combine+partial_store", "owner_poll_sum_update", "z_poll"]
    for mode in ("finito", "lfinito"):
        sw = BatchSweeper(N, 4096, 2, HostRNG(1)); idx, bp = csr(sw.take(sw.d))
        if mode == "finito":
            e.finito_init(np.ones(d), gam, hat); e.finito_steps(idx, bp)
        else:
            e.lfinito_init(np.ones(d), gam, hat); e.lfinito_outer(np.arange(1, sw.d + 1), 4096)
        out = (C.c_longlong * 8)()
        fn.argtypes = [C.c_void_p, C.c_void_p]; fn(e.h, out)
        print(f"{mode} batch 4096, cycles per batch (CTA 0): " + ", ".join(f"{n} {out[i] / sw.d:.0f}" for i, n in enumerate(names)), flush=True)

fn = getattr(e.lib, "ciao_debug_batch_trace", None)
if fn is not None and os.environ.get("CIAO_BATCH_EXCHANGE") != "barrier":   # per-CTA timeline of the middle batch (flagged-word exchange)
    import ctypes as C, json
    fn.argtypes = [C.c_void_p, C.c_void_p]
    res = {}
    for mode in ("finito", "lfinito"):
        sw = BatchSweeper(N, 4096, 2, HostRNG(1)); idx, bp = csr(sw.take(sw.d))
        if mode == "finito":
            e.finito_init(np.ones(d), gam, hat); e.finito_steps(idx, bp)
        else:
            e.lfinito_init(np.ones(d), gam, hat); e.lfinito_outer(np.arange(1, sw.d + 1), 4096)
        out = (C.c_ulonglong * 9600)(); fn(e.h, out)
        a = np.array(out[:], dtype=np.int64).reshape(1200, 8)
        a = a[a[:, 1] > 0]
        t0 = a[:, 1].min()
        rel_ns = a[:, 1:] - t0
        names = ["rows_start", "rows_end(sg0)", "owner_done", "z_received", "all_subgroups", "partial_sent", "owner_polled"]
        print(f"{mode}: {len(a)} CTAs, ns after the first CTA started the batch (min / median / max):")
        for i, n in enumerate(names):
            v = rel_ns[:, i]
            print(f"   {n:14s} {v.min():7d} {int(np.median(v)):7d} {v.max():7d}" + (f"   owners only: {v[:128].min()} {int(np.median(v[:128]))} {v[:128].max()}" if n.startswith("owner") else ""))
        dur = rel_ns[:, 1] - rel_ns[:, 0]
        print(f"   rows duration {dur.min()} / {int(np.median(dur))} / {dur.max()} ns; by SM parity (even/odd SM id) median {int(np.median(dur[a[:,0]%2==0]))} / {int(np.median(dur[a[:,0]%2==1]))}")
        slow = np.argsort(-rel_ns[:, 1])[:12]
        print("   latest rows_end: " + ", ".join(f"cta{int(i)}@sm{int(a[i,0])}:{int(rel_ns[i,1])}" for i in slow))
        res[mode] = {"smid": a[:, 0].tolist(), "ns": rel_ns.tolist()}
        fd = getattr(e.lib, "ciao_debug_batch_dur", None)
        if fd is not None:
            fd.argtypes = [C.c_void_p, C.c_void_p]
            o2 = (C.c_uint * 4800)(); fd(e.h, o2)
            du = np.array(o2[:], dtype=np.float64).reshape(1200, 4)[:len(a)]
            cc = np.corrcoef(du.T)
            print(f"   rows phase of 4 consecutive batches: per-SM correlation between batches {cc[0,1]:.2f} {cc[1,2]:.2f} {cc[2,3]:.2f} {cc[0,3]:.2f}; "
                  f"spread (max/median) {[round(float(du[:,i].max()/np.median(du[:,i])),3) for i in range(4)]}")
            res[mode]["dur"] = du.tolist()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/batch_trace.json", "w"))

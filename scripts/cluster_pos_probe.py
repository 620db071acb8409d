"""SVRG inner step time as a function of the cluster position (CIAO_SEQ_CLUSTER_POS), next to the exchange floor."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--worker" in sys.argv:
    import ciao_pkg; ciao_pkg.load()
    from ciaoalgorithms_jl_b200 import _lib as L
    from ciaoalgorithms_jl_b200.engine import Engine
    N, d = 1 << 19, 4096
    e = Engine(0)
    e.gen_synthetic(L.SYNTH_LASSO, N, d, 0x5EED0003, scale=float(N)); e.set_reg(L.REG_NORML1, N / 100.0)
    Lmax = N * e.max_row_sqnorm()
    idx = np.random.default_rng(1).integers(1, N + 1, size=1 << 18, dtype=np.int64)
    e.svrg_init(np.zeros(d), 1.0 / (7.0 * Lmax), True)
    e.svrg_epoch(idx)
    ts = []
    for _ in range(3):
        e.svrg_epoch(idx); ts.append(e.last_timing().last_seq_ms)
    print(f"pos {os.environ.get('CIAO_SEQ_CLUSTER_POS', '0'):>2}: {1e3 * min(ts) / len(idx):.4f} us/step smid {e.last_seq_placement()}", flush=True)
    if os.environ.get("CIAO_SEQ_CLUSTER_POS", "0") == "0":
        for mode in (1,):
            r = e.measure_exchange(8, 4, 100000, mode)
            print("   full-grid floor mode1 ns:", [round(x[0], 1) for x in r])
            print("   lone:", e.measure_exchange(8, 4, 100000, mode, max_clusters=1))
    e.close()
else:
    for pos in range(18):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env={**os.environ, "CIAO_SEQ_CLUSTER_POS": str(pos)}, timeout=300)

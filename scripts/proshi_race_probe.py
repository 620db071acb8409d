"""Diagnostic for the ProShI long-run bitwise test: where and how does av leave the oracle's trajectory?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ciao_pkg; ciao_pkg.load()
from oracle import oracle as orc
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine
from test_gpu_hazard_stress import stress_indices, distances

for N, n, K in ((40, 1024, 200_000), (24, 1024, 200_000), (64, 1024, 200_000), (40, 2048, 100_000)):
    idx = stress_indices(N, K, 7 * N + (n if n == 1024 else 1024))
    ptr = np.arange(K + 1, dtype=np.int64)
    ptr = ptr[np.sort(np.unique(np.concatenate([[0, K], np.random.default_rng(3).integers(0, K, size=K // 2)])))]
    Q, _ = orc.gen_rows(orc.SYN_SHARING, n, 0x5EED0005, 0, N)
    eta = 10.0 * N
    p = orc.Problem(orc.LOSS_DIAGQUAD, Q, np.ones((N, n)), box=(-2.0, 2.0), eta=eta).set_reg(orc.REG_INDBOX, lo=-np.inf, hi=np.ones(n))
    gam = 0.999 * N / (np.abs(Q).max(axis=1) + eta)
    dist = distances(idx)
    runs = []
    for rep in range(3):
        ref = orc.ProshiState(p, np.zeros(n), gam)
        with Engine(0) as e:
            e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005)
            e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
            e.proshi_init(np.zeros(n), gam, ref.hat_gamma)
            first_bad = None
            nb = len(ptr) - 1
            chunk = 2000
            for b0 in range(0, nb, chunk):
                b1 = min(nb, b0 + chunk)
                pp = ptr[b0:b1 + 1] - ptr[b0]
                ii = idx[ptr[b0]:ptr[b1]]
                ref.steps(ii, pp)
                e.proshi_steps(ii, pp)
                if first_bad is None:
                    av = e.get_vec(L.VEC_AV)
                    if not np.array_equal(av, ref.av):
                        first_bad = (b0, b1, int(ptr[b0]), int(ptr[b1]))
                        bad = np.nonzero(av != ref.av)[0]
                        print(f"N={N} n={n} rep={rep}: first divergence in batches {b0}..{b1} (steps {ptr[b0]}..{ptr[b1]}): {len(bad)} columns differ, "
                              f"cols {bad[:12].tolist()} max|diff| {np.abs(av - ref.av).max():.3e}; table equal: {np.array_equal(e.get_table_rows(), ref.s)}", flush=True)
                        seg = dist[ptr[b0]:ptr[b1]]
                        print("   distance histogram in the chunk (16..24):", np.bincount(seg, minlength=30)[14:26].tolist())
            av = e.get_vec(L.VEC_AV)
            runs.append(av.copy())
            print(f"N={N} n={n} rep={rep}: chunked run final av equal: {np.array_equal(av, ref.av)}, z equal {np.array_equal(e.get_vec(L.VEC_Z), ref.z)}, "
                  f"table equal {np.array_equal(e.get_table_rows(), ref.s)}", flush=True)
    print(f"N={N} n={n}: run-to-run identical: {all(np.array_equal(runs[0], r) for r in runs)}")
    # one call with all steps (the failing test's shape), three times
    outs = []
    for rep in range(3):
        ref = orc.ProshiState(p, np.zeros(n), gam)
        ref.steps(idx, ptr)
        with Engine(0) as e:
            e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005)
            e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
            e.proshi_init(np.zeros(n), gam, ref.hat_gamma)
            e.proshi_steps(idx, ptr)
            av = e.get_vec(L.VEC_AV)
            outs.append(av.copy())
            bad = np.nonzero(av != ref.av)[0]
            print(f"N={N} n={n} single call rep={rep}: {len(bad)} columns differ {bad[:16].tolist()} max|diff| {np.abs(av - ref.av).max():.3e}", flush=True)
    print(f"N={N} n={n}: single-call run-to-run identical: {all(np.array_equal(outs[0], r) for r in outs)}")

"""Latency floor of the cluster exchange (ciao_measure_exchange, seq_floor.cu): per cluster placement, modes 0/1."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200.engine import Engine
out = {}
with Engine(0) as e:
    for cluster, warps in ((8, 4), (8, 2), (4, 8), (4, 4), (16, 2), (2, 8)):
        for mode in (0, 1):
            e.measure_exchange(cluster, warps, 2000, mode)           # warm
            r = e.measure_exchange(cluster, warps, 200000, mode)
            ns = np.array([x[0] for x in r]); cyc = np.array([x[1] for x in r])
            key = f"C{cluster}_W{warps}_mode{mode}"
            out[key] = {"clusters": len(r), "ns_min": float(ns.min()), "ns_median": float(np.median(ns)), "ns_max": float(ns.max()),
                        "cyc_min": float(cyc.min()), "cyc_median": float(np.median(cyc)), "cyc_max": float(cyc.max())}
            print(key, json.dumps(out[key]), flush=True)
            if cluster == 8 and warps == 4:
                for k, (a, b, sm) in enumerate(r):
                    print(f"   cluster {k:2d}: {a:7.1f} ns {b:7.1f} cyc  smid {sm}")
    # run it twice more alone to see whether the placement → latency map is stable launch to launch
    for rep in range(2):
        r = e.measure_exchange(8, 4, 200000, 1)
        print("repeat", rep, [round(x[0], 1) for x in r], [x[2][0] for x in r])
    r1 = e.measure_exchange(8, 4, 200000, 1, max_clusters=1)
    print("single cluster launch:", r1)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "exchange_floor_r2.json"), "w"), indent=1)

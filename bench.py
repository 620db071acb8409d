#!/usr/bin/env python
"""bench.py — headline benchmark of the CIAOAlgorithms.jl hot path on B200.

Workload (BASELINE.json configs[2], "C3"): Lasso N = 2^22, d = 4096, fp64, SVRG++
(γ = 1/(7·L_max), m0 = N/16 doubling per outer iteration, schedule restarted every 5).
One "step" = one SVRG++ outer iteration = ciao_svrg_epoch: the persistent inner-epoch
kernel over m host-sampled rows + the full-gradient streaming pass over all N rows
(SVRG_basic.jl:71-96).  metric = epochs/s with 1 epoch = N component-gradient
evaluations in one-data-pass accounting: a step with inner length m is (m + N)/N epochs
(SURVEY.md §8d).

  value     : indices already resident in HBM, no read-back — device time (CUDA events)
  e2e       : the public API (solvers.iterator → next(state) → solution(state)): host RNG draw,
              H2D of the index sequence from pinned memory and D2H of the solution inside the
              timed region.  The data matrix is uploaded/generated once per solve (like F in
              `solver(x0; F=...)`), reported under "setup".
  roofline  : the full-gradient pass kernel (row_pass_kernel), HBM-bound; algorithmic bytes
              = N·(d_pad+8)·8 per launch (the row records incl. the 8-scalar tail b_i, λ_i, …).
  N > 1     : one process per GPU.  96 % of a step is the sequential inner epoch, which does not shard
              (step k+1 depends on step k: "replicas only"), so the default workload at
              N > 1 is N independent SVRG++ solves of the C3 problem — one per GPU, each with its own
              index stream, no collective on the data path → "weak" scaling; `value` = the epochs all
              ranks processed ÷ the max-over-ranks time.  The part of the path that does shard — the
              full-gradient pass, row-windowed over the ranks + NCCL allreduce of the d-vector — is
              timed in the same run after the solves and reported under "full_gradient_sharded".
              `--workload svrgpp-strong` runs ONE solve on N GPUs (pass sharded, inner epoch replicated;
              Amdahl-bound), `--workload fullgrad` the sharded pass alone with 2^22 rows per GPU (weak),
              `--workload saga-init` the sharded table-init pass, `--workload svrgpp-sharded` a solve
              over row shards with remote rows fetched over NVLink.
  --impl reference : the CPU restatement of the reference (oracle/, kind "port"; Julia is not
              installed) on a bounded sample of the same workload, single thread like the reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_DATA, SEED_IDX = 0x5EED0003, 0x1D0003
SCHEDULE = 5  # m = m0·2^(k mod 5): N/16 … N


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="svrgpp", choices=["svrgpp", "svrgpp-strong", "fullgrad", "svrgpp-sharded", "saga-init"])
    ap.add_argument("--rows-log2", type=int, default=22)
    ap.add_argument("--d", type=int, default=4096)
    ap.add_argument("--cpu-rows-log2", type=int, default=17)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the c2/c5/c4 blocks measured after the headline")
    ap.add_argument("--tune", default="", help="pass_threads,pass_stages,pass_ctas,seq_cluster,seq_threads")
    return ap.parse_args()


def m_of(step, N):
    return (N // 16) << (step % SCHEDULE)


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md), read through NVML in-process:
    forking `nvidia-smi` every 200 ms from a process that maps 137 GB stalls the launching thread
    (measured: ≈ 190 ms over a 5 s region), so nvidia-smi is only the fallback."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu]) if vis and vis.split(",")[gpu].isdigit() else gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _sample(self):
        if self.nv is not None:
            nv = self.nv
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            names = [n for n, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                                      ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                                      ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                                      ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)) if r & bit]
            return float(sm), float(mx), names
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        out = [c.strip() for c in out]
        names = [n for n, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:6])
                 if v.lower().startswith("active")]
        return float(out[0]), float(out[1]), names

    def run(self):
        while not self.stop_flag:
            try:
                self.rows.append(self._sample())
            except Exception:
                pass
            time.sleep(0.1 if self.nv is not None else 1.0)

    def summary(self):
        self.stop_flag = True
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = sorted({n for r in self.rows for n in r[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nv is not None else "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
def _cpu_run(Ns, d, steps, warmup, seed_idx, schedule=True):
    """`steps` SVRG++ outer iterations of the oracle on the first Ns rows of the C3 generator; per-phase seconds."""
    from oracle import oracle as orc
    A, b = orc.gen_rows(orc.SYN_LASSO, d, SEED_DATA, 0, Ns)
    p = orc.Problem(orc.LOSS_LS, A, b, np.full(Ns, float(Ns))).set_reg(orc.REG_NORML1, lam=Ns / 100.0)
    gamma = 1 / (7 * Ns * p.max_row_sqnorm())
    st = orc.SVRGState(p, np.zeros(d), gamma, m=Ns // 16, plus=True)
    rng = np.random.default_rng(seed_idx)
    evals, inner_steps, t_inner, t_pass = 0, 0, 0.0, 0.0
    for k in range(-warmup, steps):
        m = (Ns // 16) << ((k % SCHEDULE) if (k >= 0 and schedule) else 0)
        idx = rng.integers(1, Ns + 1, size=m, dtype=np.int64)
        st.m = m
        t0 = time.perf_counter()
        st.inner(idx)
        t1 = time.perf_counter()
        st.av[:] = p.full_gradient(st.z_full, 1.0 / Ns)
        t2 = time.perf_counter()
        if k >= 0:
            evals += m + Ns
            inner_steps += m
            t_inner += t1 - t0
            t_pass += t2 - t1
    return {"rows": Ns, "steps": steps, "evals": evals, "seconds": t_inner + t_pass, "us_per_inner_step": 1e6 * t_inner / inner_steps,
            "us_per_pass_row": 1e6 * t_pass / (steps * Ns), "us_per_component_gradient": 1e6 * (t_inner + t_pass) / evals}, p, st


def cpu_reference(args, N_full, d, steps, warmup, seed_idx=SEED_IDX, extras=True):
    """The reference's loop (oracle port, one thread like the reference) on a bounded sample: `steps` SVRG++ outer iterations
    with the bench's own m schedule on 2^cpu_rows_log2 rows of the same generator, extrapolated per component gradient to
    N_full.  With `extras`: the same per-gradient costs at a quarter and at four times the sample (is the extrapolation
    linear?) and the separately labelled all-cores full-gradient pass SURVEY.md §8d allows."""
    Ns = 1 << args.cpu_rows_log2
    main, p, st = _cpu_run(Ns, d, steps, warmup, seed_idx)
    per_eval = main["seconds"] / main["evals"]
    cb = {"value": 1.0 / (per_eval * N_full), "unit": "epochs/s", "cores": 1, "kind": "port",
          "sample": f"oracle (C restatement of the single-threaded reference; Julia is not installed) SVRG++ outer "
                    f"iterations on {Ns} rows x {d} (same generator), {steps} steps of the bench's m schedule, {main['evals']} "
                    f"component gradients in {main['seconds']:.2f} s on 1 core (the reference has no threading; its BLAS calls "
                    f"are 1 x d gemv); extrapolated per component gradient to N = {N_full}",
          "us_per_inner_step": main["us_per_inner_step"], "us_per_pass_row": main["us_per_pass_row"], "seconds": main["seconds"]}
    if extras:
        from oracle import oracle as orc
        cores = orc.num_threads()
        x = st.z_full.copy()
        p.full_gradient_omp(x, 1.0 / Ns, cores)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            p.full_gradient_omp(x, 1.0 / Ns, cores)
        t_omp = (time.perf_counter() - t0) / reps
        cb["all_cores_full_gradient"] = {
            "what": "NOT the reference's algorithm (it is single-threaded): the same per-row operation sequence with the rows split "
                    "over all host cores, labelled separately as SURVEY.md 8d allows", "cores": cores, "rows": Ns,
            "us_per_row": 1e6 * t_omp / Ns, "extrapolated_pass_s_at_N": t_omp / Ns * N_full, "gbs": Ns * d * 8 / t_omp / 1e9}
        del p, st
        lin = [main]
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 0
        for lg, st_n, sched in ((args.cpu_rows_log2 - 2, 5, True), (args.cpu_rows_log2 + 2, 1, False)):
            need = (1 << lg) * d * 8
            if need * 2.5 > avail:
                lin.append({"rows": 1 << lg, "skipped": f"needs {need / 2**30:.0f} GiB of host memory"})
                continue
            r, p2, st2 = _cpu_run(1 << lg, d, st_n, 0, seed_idx, schedule=sched)
            del p2, st2
            lin.append(r)
        ok = [r for r in lin if "skipped" not in r]
        cb["linearity"] = {"what": "per-gradient cost of the port at three sample sizes (the extrapolation to N assumes it is flat)",
                           "runs": sorted(lin, key=lambda r: r["rows"]),
                           "max_dev_us_per_pass_row": max(abs(r["us_per_pass_row"] / main["us_per_pass_row"] - 1) for r in ok),
                           "max_dev_us_per_inner_step": max(abs(r["us_per_inner_step"] / main["us_per_inner_step"] - 1) for r in ok)}
    return cb, main["seconds"] / steps


def _reference_replica(job):
    args, N, d, steps, warmup, r = job
    return cpu_reference(args, N, d, steps, warmup, seed_idx=SEED_IDX + r, extras=False)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, d = 1 << args.rows_log2, args.d
    steps = max(1, args.steps)          # the same number of outer iterations, on the same m schedule, as the GPU arm times
    G = args.gpus if args.workload == "svrgpp" else 1
    if G > 1:
        # the GPU arm at N > 1 runs G independent solves, one per GPU: the CPU arm runs the same G solves concurrently,
        # one single-threaded reference loop per host core (the reference itself has no threading), 2 GiB of rows each
        import multiprocessing as mp
        args.cpu_rows_log2 = min(args.cpu_rows_log2, 16)
        with mp.get_context("spawn").Pool(G) as pool:
            res = pool.map(_reference_replica, [(args, N, d, steps, min(args.warmup, 2), r) for r in range(G)])
        s_per_step = max(r[1] for r in res)                    # the job ends with its slowest replica
        cb = dict(res[0][0])
        cb["value"] = G * min(r[0]["value"] for r in res)      # whole job: G solves at the pace of the slowest
        cb["cores"] = G
        cb["sample"] = f"{G} concurrent replicas, one per host core, each: " + cb["sample"]
    else:
        cb, s_per_step = cpu_reference(args, N, d, steps, min(args.warmup, 2))
    line = {"impl": "reference", "metric": "epochs/s (Lasso 4M x 4096 fp64, SVRG++)", "value": cb["value"], "unit": "epochs/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * s_per_step,
            "higher_is_better": True, "scaling": "strong" if args.workload == "svrgpp-strong" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"C3 Lasso N=2^{args.rows_log2} d={d} fp64 SVRG++ gamma=1/(7 L_max) m=N/16*2^(k mod 5): persistent inner epoch + "
                                    f"full-gradient pass" + (f"; {G} independent solves" if G > 1 else "")),
                       "sample": f"the reference's loop (CPU restatement) on 2^{args.cpu_rows_log2} rows of the same generator, same schedule, "
                                 f"extrapolated per component gradient to N=2^{args.rows_log2}",
                       "epoch": "N component-gradient evaluations; step = (m + N)/N epochs", "seeds": [SEED_DATA, SEED_IDX]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# The other BASELINE.json configurations, measured in the same run and printed under their own keys:
# none of them is the headline `value`; each carries its own roofline figure against the measured copy peak.
def bench_c2(local, peak, K=2):
    """configs[1] "C2": L1-logistic N = 2^20, d = 1024 — SAGA + Finito steps (batch 1, latency-bound), their table-init passes
    and the static-minibatch passes (HBM-bound).  SAGA_basic.jl:41-68, Finito_basic.jl:76-121, Finito_LFinito.jl:78-103."""
    from ciaoalgorithms_jl_b200 import _lib as L
    from ciaoalgorithms_jl_b200.engine import Engine
    from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, csr
    N, d = 1 << 20, 1024
    ld = d + 8
    out = {"workload": "C2 L1-logistic N=2^20 d=1024 fp64, NormL1(1/N), x0 = ones", "N": N, "d": d}
    with Engine(local) as e:
        e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0)
        e.set_reg(L.REG_NORML1, 1.0 / N)
        Lmax = 0.25 * e.max_row_sqnorm()
        x0 = np.ones(d)
        rng = HostRNG(0x1D0002)
        f0 = sum(e.objective(x0))

        def pass_stats():
            t = e.last_timing()
            g = t.last_pass_bytes / t.last_pass_ms / 1e6
            return {"ms": t.last_pass_ms, "GBs": g, "frac_of_peak": g / peak, "algorithmic_bytes": int(t.last_pass_bytes)}

        e.saga_init(x0, 1 / (3 * Lmax), False)
        e.saga_init(x0, 1 / (3 * Lmax), False)
        out["saga_table_init"] = pass_stats()
        e.saga_steps(rng.rand_vec(N, N // 8))                       # warm-up
        ms = 0.0
        for _ in range(K):
            e.saga_steps(rng.rand_vec(N, N))
            ms += e.last_timing().last_seq_ms
        out["saga"] = {"epochs": K, "us_per_step": 1e3 * ms / (K * N), "epochs_per_s": K / (ms / 1e3), "bytes_per_step": 24 * d + 64,
                       "objective_start": f0, "objective": sum(e.objective(e.get_vec(L.VEC_Z))), "smids": e.last_seq_placement()}
        gam = np.full(N, 0.999 * N / Lmax)
        hat = 1 / np.sum(1 / gam)
        e.finito_init(x0, gam, hat)
        out["finito_table_init"] = pass_stats()
        bp1 = np.arange(N + 1, dtype=np.int64)
        e.finito_steps(rng.rand_vec(N, N // 8), bp1[: N // 8 + 1])
        ms = 0.0
        for _ in range(K):
            e.finito_steps(rng.rand_vec(N, N), bp1)
            ms += e.last_timing().last_seq_ms
        out["finito"] = {"epochs": K, "us_per_step": 1e3 * ms / (K * N), "epochs_per_s": K / (ms / 1e3), "sweeping": 1,
                         "objective": sum(e.objective(e.get_vec(L.VEC_Z)))}
        for r in (4096, 65536):                                     # static minibatches: streaming passes (batch.cu)
            e.finito_init(x0, gam, hat)
            sw = BatchSweeper(N, r, 2, rng)
            idx, bp = csr(sw.take(sw.d))
            e.finito_steps(idx, bp)
            ms = 0.0
            for _ in range(K):
                idx, bp = csr(sw.take(sw.d))
                e.finito_steps(idx, bp)
                ms += e.last_timing().last_seq_ms
            g = K * N * (ld + 2 * d) * 8 / ms / 1e6
            out[f"finito_batch{r}"] = {"epochs": K, "epochs_per_s": K / (ms / 1e3), "GBs": g, "frac_of_peak": g / peak,
                                       "us_per_batch": 1e3 * ms / (K * sw.d), "objective": sum(e.objective(e.get_vec(L.VEC_Z)))}
            e.lfinito_init(x0, gam, hat)
            order = np.arange(1, sw.d + 1, dtype=np.int64)
            e.lfinito_outer(order, r)
            ms = 0.0
            for _ in range(K):
                e.lfinito_outer(order, r)
                ms += e.last_timing().last_seq_ms
            g = K * N * ld * 8 / ms / 1e6
            out[f"lfinito_sweep_batch{r}"] = {"sweeps": K, "sweeps_per_s": K / (ms / 1e3), "GBs": g, "frac_of_peak": g / peak,
                                              "us_per_batch": 1e3 * ms / (K * sw.d), "objective": sum(e.objective(e.get_vec(L.VEC_Z)))}
    return out


def bench_c5(local, peak, K=3):
    """configs[4] "C5": sharing problem, N = 2^18 blocks of n = 1024 — ProShI block steps (batch 1: latency-bound chain per
    column; batch 4096: HBM-bound), table init and the in-place solution.  ProShI_basic.jl:76-132."""
    from ciaoalgorithms_jl_b200 import _lib as L
    from ciaoalgorithms_jl_b200.engine import Engine
    from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, csr
    N, n = 1 << 18, 1024
    out = {"workload": "C5 sharing N=2^18 blocks x n=1024 fp64: diag Quadratic + SqrDistL2(box) blocks, g = IndBox(-inf, 1)", "N": N, "n": n}
    with Engine(local) as e:
        e.gen_synthetic(L.SYNTH_SHARING, N, n, 0x5EED0005)
        e.set_reg(L.REG_INDBOX, -np.inf, np.ones(n))
        gam = 0.999 * N / np.full(N, 10.0 + 10.0 * N)
        e.proshi_init(np.zeros(n), gam, float(np.sum(gam)))
        e.proshi_init(np.zeros(n), gam, float(np.sum(gam)))
        t = e.last_timing()
        g = t.last_pass_bytes / t.last_pass_ms / 1e6
        out["table_init"] = {"ms": t.last_pass_ms, "GBs": g, "frac_of_peak": g / peak}
        rng = HostRNG(0x1D0005)
        for r in (1, 4096):
            sw = BatchSweeper(N, r, 2, rng)
            idx, bp = csr(sw.take(sw.d))
            e.proshi_steps(idx, bp)
            ms = 0.0
            for _ in range(K):
                idx, bp = csr(sw.take(sw.d))
                e.proshi_steps(idx, bp)
                ms += e.last_timing().last_seq_ms
            ga, gd = 24.0 * n * N * K / ms / 1e6, 32.0 * n * N * K / ms / 1e6
            out[f"proshi_batch{r}"] = {"sweeps": K, "us_per_block": 1e3 * ms / (K * N), "sweeps_per_s": K / (ms / 1e3),
                                       "GBs_algorithmic_24n": ga, "GBs_dram_32n": gd, "frac_of_peak_dram": gd / peak}
        e.proshi_solution(None)
        e.proshi_solution(None)
        t = e.last_timing()
        g = t.last_pass_bytes / t.last_pass_ms / 1e6
        out["solution_inplace"] = {"ms": t.last_pass_ms, "GBs": g, "frac_of_peak": g / peak}
        out["sum_x_first3"] = e.table_colsum()[:3].tolist()
    return out


def bench_c4(local, rank, world, peak, d, K, W, init_comm, barrier, max_over_ranks, same_on_all_ranks):
    """configs[3] "C4": Lasso rows sharded over the GPUs — the full-gradient pass (SVRG_basic.jl:58-63 over a sharded F) with
    min(2^22, 2^24/G) rows per GPU (N = 2^24 in total at 4 and 8 GPUs) and the SAGA table-init pass (SAGA_basic.jl:41-47) with
    min(2^21, 2^24/G) rows + their table rows per GPU, each closed by the exchange of the d-vector.  Weak scaling."""
    from ciaoalgorithms_jl_b200 import _lib as L
    from ciaoalgorithms_jl_b200.engine import Engine
    out = {}
    ld = (d + 3) // 4 * 4 + 8
    for what, rows in (("full_gradient", min(1 << 22, (1 << 24) // world)), ("saga_table_init", min(1 << 21, (1 << 24) // world))):
        N = rows * world
        with Engine(local) as e:
            e.gen_synthetic(L.SYNTH_LASSO, N, d, SEED_DATA, scale=float(N), row0=rank * rows, n_rows=rows)
            e.set_reg(L.REG_NORML1, N / 100.0)
            init_comm(e)
            x = np.full(d, 1e-3)
            e.set_vec(L.VEC_X, x)
            gamma = 1e-9

            def one(read=False):
                if what == "full_gradient":
                    return e.full_gradient(None, 1.0 / N, out=read)
                e.saga_init(x, gamma, False)
                return e.get_vec(L.VEC_AV) if read else None

            for _ in range(W):
                one()
            barrier()
            e.timer_begin()
            for _ in range(K):                  # timed: back-to-back passes
                one()
            ms = max_over_ranks(e.timer_end()) / K
            barrier()
            kms, tms = [], []
            for _ in range(K):                  # per-kernel device times (reading them synchronises: separate loop)
                one()
                tm = e.last_timing()
                kms.append(tm.last_pass_ms)
                tms.append(tm.last_tail_ms)
            barrier()
            kernel_ms = max_over_ranks(float(np.mean(kms)))
            res = one(read=True)
            bytes_per_gpu = rows * (ld + (d if what == "saga_table_init" else 0)) * 8
            out[what] = {"N_total": N, "rows_per_gpu": rows, "ms_per_pass": ms, "kernel_ms": kernel_ms, "tail_us": 1e3 * (ms - kernel_ms),
                         "tail_kernel_us": 1e3 * max_over_ranks(float(np.mean(tms))),
                         "aggregate_gbs": world * bytes_per_gpu / ms / 1e6, "kernel_gbs_per_gpu": bytes_per_gpu / kernel_ms / 1e6,
                         "frac_of_peak_per_gpu": bytes_per_gpu / ms / 1e6 / peak, "epochs_per_s_2p22_rows": world * (rows / float(1 << 22)) / (ms / 1e3),
                         "bitwise_equal_across_ranks": bool(same_on_all_ranks(res)), "checksum": float(np.sum(res)), "passes": K}
            barrier()
    return out


def bench_sharded_minibatch(local, rank, world, init_comm, barrier, max_over_ranks, same_on_all_ranks):
    """configs[1] "C2" rows as interleaved shards (blocks of 256 rows dealt round-robin to the GPUs, the N x d table lives with its
    rows): one epoch of static Finito minibatches (Finito_basic.jl:110-118) and one LFinito sweep (Finito_LFinito.jl:91-100), each
    a single persistent kernel per rank whose column owners exchange the ranks' sums through the peer arenas (batch.cu).  Batches
    of 4096 rows per GPU (weak) and of 65 536 rows in total (strong); compare with C2's single-GPU µs per batch."""
    from ciaoalgorithms_jl_b200 import _lib as L
    from ciaoalgorithms_jl_b200.engine import Engine
    from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, csr, interleaved_rows
    N, d, B = 1 << 20, 1024, 256
    n_loc = len(interleaved_rows(N, B, world, rank))
    out = {"N": N, "d": d, "layout": f"interleaved, blocks of {B} rows", "rows_per_gpu": n_loc}
    with Engine(local) as e:
        e.set_row_interleave(B, rank, world)
        e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0, row0=rank * B, n_rows=n_loc)
        e.set_reg(L.REG_NORML1, 1.0 / N)
        init_comm(e)
        gam = np.full(N, 0.999 * N / (0.25 * e.max_row_sqnorm()))
        hat = 1 / np.sum(1 / gam)
        for r in (4096 * world, 65536):
            sw = BatchSweeper(N, r, 2, HostRNG(1))
            idx, bp = csr(sw.take(sw.d))
            e.finito_init(np.ones(d), gam, hat)
            e.finito_steps(idx, bp)
            barrier()
            e.finito_steps(idx, bp)
            tf = max_over_ranks(e.last_timing().last_seq_ms)
            zf = e.get_vec(L.VEC_Z)
            e.lfinito_init(np.ones(d), gam, hat)
            order = np.arange(1, sw.d + 1)
            e.lfinito_outer(order, r)
            barrier()
            e.lfinito_outer(order, r)
            tl = max_over_ranks(e.last_timing().last_seq_ms)
            zl = e.get_vec(L.VEC_Z)
            out[f"batch{r}"] = {"rows_per_gpu_and_batch": r // world, "finito_us_per_batch": 1e3 * tf / sw.d, "finito_epochs_per_s": 1e3 / tf,
                                "lfinito_us_per_batch": 1e3 * tl / sw.d, "lfinito_sweeps_per_s": 1e3 / tl,
                                "z_bitwise_equal_across_ranks": bool(same_on_all_ranks(zf)) and bool(same_on_all_ranks(zl)),
                                "objective": sum(e.objective(zl))}
            barrier()
    return out


def bench_host_rows(local):
    """F handed over from HOST memory (ciao_set_rows, what `solver(x0; F=...)` does with a Julia F): seconds and GB/s of the one big
    copy that the e2e figure amortises over the solve.  2^20 x 4096 (34 GB) when the host has the memory, else smaller."""
    from ciaoalgorithms_jl_b200 import _lib as L
    from ciaoalgorithms_jl_b200.engine import Engine
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 8 << 30
    d = 4096
    lg = 20
    while lg > 14 and (1 << lg) * d * 8 * 1.3 > avail:
        lg -= 1
    n = 1 << lg
    blk, rhs = Engine.gen_host(L.SYNTH_LASSO, d, SEED_DATA, 0, 1 << 12)
    A = np.empty((n, d))
    for r0 in range(0, n, 1 << 12):
        A[r0:r0 + (1 << 12)] = blk
    b = np.tile(rhs, n >> 12)
    out = {"rows": n, "d": d, "bytes": int(A.nbytes)}
    with Engine(local) as e:
        t0 = time.perf_counter()
        e.set_rows(L.LOSS_LS, A, b, float(n))
        e.sync()
        dt = time.perf_counter() - t0
        out.update({"seconds": dt, "GBs": A.nbytes / dt / 1e9, "source": "pageable host memory (numpy)"})
        x = np.full(d, 1e-3)
        g1 = e.full_gradient(x, 1.0 / n)
        out["checksum"] = float(np.sum(g1))
    return out


# ------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import ciao_pkg
    ciao_pkg.load()
    from ciaoalgorithms_jl_b200 import _lib as L
    from ciaoalgorithms_jl_b200 import operators as ops
    from ciaoalgorithms_jl_b200 import solvers
    from ciaoalgorithms_jl_b200.engine import Engine
    from ciaoalgorithms_jl_b200.sampling import HostRNG

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libciao_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    table_init = args.workload == "saga-init"   # sharded K2: SAGA table init pass (read A, write the N×d table) + allreduce
    if table_init and args.rows_log2 == 22:
        args.rows_log2 = 21                      # rows + table of a shard must fit: 2^21 × 4096 × 8 B × 2 = 137 GB per GPU (C4 on 8 GPUs)
    rows_per_gpu = 1 << args.rows_log2
    d = args.d
    weak_pass = args.workload == "fullgrad" or table_init
    sharded = args.workload == "svrgpp-sharded"      # rows sharded over the GPUs, inner epoch reads remote rows over NVLink
    strong = args.workload == "svrgpp-strong" and world > 1    # ONE solve on all GPUs: pass row-windowed, inner epoch replicated
    replicas = args.workload == "svrgpp" and world > 1         # one independent solve per GPU (the inner epoch does not shard)
    N = rows_per_gpu * world if (weak_pass or sharded) else rows_per_gpu
    e = Engine(local)
    if args.tune:
        e.set_tuning(*[int(v) for v in args.tune.split(",")])
    t0 = time.perf_counter()
    if weak_pass or sharded:
        e.gen_synthetic(L.SYNTH_LASSO, N, d, SEED_DATA, scale=float(N), row0=rank * rows_per_gpu, n_rows=rows_per_gpu)
    else:
        e.gen_synthetic(L.SYNTH_LASSO, N, d, SEED_DATA, scale=float(N))
    e.set_reg(L.REG_NORML1, N / 100.0)
    e.sync()
    setup_s = time.perf_counter() - t0
    def init_comm(eng=None):
        """NCCL communicator (its all-gather carries the per-row step scalars of a sharded solve) + the one-shot peer-memory
        exchange that the passes' all-reduce uses (ciao_comm_p2p_*: deterministic rank-ordered sum in the pass's tail kernel)."""
        eng = e if eng is None else eng
        obj = [Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, 0)
        eng.comm_init(obj[0], rank, world)
        if os.environ.get("CIAO_BENCH_EXCHANGE", "p2p") == "p2p":
            hs = [None] * world
            dist.all_gather_object(hs, eng.comm_p2p_handle())
            eng.comm_p2p_attach(rank, world, hs)

    def same_on_all_ranks(a):
        blobs = [None] * world
        dist.all_gather_object(blobs, np.ascontiguousarray(a).tobytes())
        return all(b == blobs[0] for b in blobs)

    if world > 1 and not replicas:       # replicas: the solves run without a communicator (a pass with one all-reduces)
        init_comm()
        if sharded:
            handles = [None] * world
            dist.all_gather_object(handles, e.rows_ipc_handle())
            e.attach_peer_rows(handles, [r * rows_per_gpu for r in range(world)], [rows_per_gpu] * world, rank)
        elif strong:
            lo, hi = (rank * N) // world, ((rank + 1) * N) // world
            e.set_pass_window(lo, hi - lo)
    gamma = 1.0 / (7.0 * N * e.max_row_sqnorm())
    x0 = np.zeros(d)
    peak, peak_src = measured_peak()
    K, W = args.steps, max(args.warmup, 3)
    ld = (d + 3) // 4 * 4 + 8

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    if weak_pass:
        # ---- sharded full-gradient pass alone (C4-style, weak scaling) -------------------------
        e.set_vec(L.VEC_X, np.full(d, 1e-3))
        x_init = np.full(d, 1e-3)

        def one_pass(x_host=None, out=False):
            if table_init:                         # SAGA_basic.jl:41-48: s_i = ∇f_i(x0) for my rows, av = Σ s_i / N (allreduce), z
                e.saga_init(x_init if x_host is None else x_host, gamma, False)
                return e.get_vec(L.VEC_AV) if out else None
            return e.full_gradient(x_host, 1.0 / N, out=out)

        for _ in range(W):
            one_pass()
        barrier()
        l0 = e.last_timing().launches
        pass_ms_list = []
        e.timer_begin()
        for _ in range(K):
            one_pass()
            pass_ms_list.append(e.last_timing().last_pass_ms)
        ms = max_over_ranks(e.timer_end())
        barrier()
        tm = e.last_timing()
        launches = tm.launches - l0
        pass_ms = max_over_ranks(float(np.mean(pass_ms_list)))
        value = K * world * (rows_per_gpu / float(1 << 22)) / (ms / 1e3)          # epochs of 2^22 rows per second, whole job
        out = one_pass(out=True)
        xh = torch.full((d,), 1e-3, dtype=torch.float64).pin_memory().numpy()
        barrier()
        e2e_t0 = time.perf_counter()
        e.timer_begin()
        for _ in range(K):
            out = one_pass(xh, out=True)
        e2e_ms = max_over_ranks(e.timer_end())
        e2e = {"value": K * world * (rows_per_gpu / float(1 << 22)) / (e2e_ms / 1e3), "unit": "epochs/s", "h2d_bytes_per_step": 8 * d, "d2h_bytes_per_step": 8 * d,
               "wall_ms": 1e3 * (time.perf_counter() - e2e_t0)}
        algo_bytes = rows_per_gpu * (ld + (d if table_init else 0)) * 8
        extra = {"table_init" if table_init else "full_gradient": {"rows_per_gpu": rows_per_gpu, "kernel_ms": pass_ms, "gbs_per_gpu": algo_bytes / pass_ms / 1e6,
                                   "aggregate_gbs": world * algo_bytes / (ms / K) / 1e6, "checksum": float(np.sum(out))}}
        scaling = "weak"
        workload = (f"C4-style sharded SAGA table init (read rows, write the N x d table), Lasso 2^{args.rows_log2} rows/GPU x d={d}, NCCL allreduce of the d-vector"
                    if table_init else
                    f"C4-style sharded full-gradient pass, Lasso 2^{args.rows_log2} rows/GPU x d={d}, NCCL allreduce of the d-vector")
        epochs_total = K * world
    else:
        # ---- SVRG++ outer iterations (C3) ---------------------------------------------------------
        seed_idx = SEED_IDX + (rank if replicas else 0)      # replicas: every GPU walks its own index stream
        rng = np.random.default_rng(seed_idx)
        idx_host = [rng.integers(1, N + 1, size=m_of(k, N), dtype=np.int64) for k in range(K)]
        idx_warm = rng.integers(1, N + 1, size=N // 16, dtype=np.int64)
        idx_dev = [torch.from_numpy(a).cuda() for a in idx_host]
        warm_dev = torch.from_numpy(idx_warm).cuda()
        torch.cuda.synchronize()
        e.svrg_init(x0, gamma, True)
        for _ in range(W):
            e.svrg_epoch(warm_dev.data_ptr(), N // 16)
        e.svrg_init(x0, gamma, True)
        barrier()
        l0 = e.last_timing().launches
        pass_ms_list, seq_ms_list, seq_mhz_list = [], [], []
        e.timer_begin()
        for k in range(K):
            e.svrg_epoch(idx_dev[k].data_ptr(), m_of(k, N))
            tm = e.last_timing()
            pass_ms_list.append(tm.last_pass_ms)
            seq_ms_list.append(tm.last_seq_ms)
            seq_mhz_list.append(e.last_seq_clock_mhz())
        ms = max_over_ranks(e.timer_end())
        barrier()
        launches = e.last_timing().launches - l0
        epoch_rows = rows_per_gpu if sharded else N      # sharded: an "epoch" stays 2^22 component gradients
        epochs_total = sum((m_of(k, N) + N) / epoch_rows for k in range(K)) * (world if replicas else 1)
        value = epochs_total / (ms / 1e3)
        pass_ms = max_over_ranks(float(np.mean(pass_ms_list)))
        inner_steps = sum(m_of(k, N) for k in range(K))
        x_dev_path = e.get_vec(L.VEC_Z_FULL)
        smids = e.last_seq_placement()
        # latency floor of the cluster exchange, measured on this GPU with the kernel's own st.async → mbarrier path (seq_floor.cu):
        # every 8-CTA cluster position reports its round time; the position the inner kernel ran on is looked up by its SM ids
        fl = e.measure_exchange(8, 4, 100000, 1)
        fl0 = e.measure_exchange(8, 4, 100000, 0)
        lone1 = e.measure_exchange(8, 4, 100000, 1, max_clusters=1)[0]     # one cluster alone on the GPU, like the inner kernel
        lone0 = e.measure_exchange(8, 4, 100000, 0, max_clusters=1)[0]
        here = [k for k, r in enumerate(fl) if sorted(r[2]) == sorted(smids)]
        floor = {"what": "ns per round of exchange + partial sum + dependent next message (mode 1) / exchange only (mode 0): `lone` = "
                         "one cluster alone on the GPU (the inner kernel's situation; same SMs when lone_same_sms), the others = all "
                         "cluster positions of the GPU running at once (clusters that share an SM slow each other down)",
                 "lone_mode1_ns": lone1[0], "lone_mode0_ns": lone0[0], "lone_mode1_cycles": lone1[1],
                 "lone_same_sms": sorted(lone1[2]) == sorted(smids),
                 "mode1_ns_min": min(r[0] for r in fl), "mode1_ns_median": float(np.median([r[0] for r in fl])),
                 "mode1_ns_max": max(r[0] for r in fl), "mode0_ns_min": min(r[0] for r in fl0),
                 "mode0_ns_median": float(np.median([r[0] for r in fl0])), "mode0_ns_max": max(r[0] for r in fl0),
                 "here_mode1_ns": fl[here[0]][0] if here else None, "here_mode0_ns": fl0[here[0]][0] if here else None,
                 "cycles_mode1_median": float(np.median([r[1] for r in fl]))}
        f_end = sum(e.objective(x_dev_path))
        f_start = sum(e.objective(x0))
        del idx_dev
        # ---- e2e through the public API: iterator protocol, host RNG, pinned H2D, D2H of the solution
        pin = torch.empty(N, dtype=torch.int64).pin_memory().numpy()

        class PinnedRNG(HostRNG):
            def rand_vec(self, n_, m_):
                pin[:m_] = self.g.integers(1, n_ + 1, size=m_, dtype=np.int64)
                return pin[:m_]

        solver = solvers.SVRG(gamma=gamma, m=N // 16, plus=True)
        it = iter(solvers.iterator(solver, x0, F=solvers.DeviceProblem(e), g=ops.NormL1(N / 100.0), N=N, rng=PinnedRNG(seed_idx)))
        state = next(it)
        barrier()
        w0 = time.perf_counter()
        e.timer_begin()
        for k in range(K):
            state.m = m_of(k, N)
            state = next(it)
            xs = solvers.solution(state)
        e2e_ms = max_over_ranks(e.timer_end())
        wall_ms = 1e3 * (time.perf_counter() - w0)
        e2e_ms = max(e2e_ms, max_over_ranks(wall_ms))     # host RNG time is outside the stream: take the wall clock
        e2e = {"value": epochs_total / (e2e_ms / 1e3), "unit": "epochs/s",
               "h2d_bytes_per_step": int(8 * inner_steps / K), "d2h_bytes_per_step": 8 * d, "wall_ms": wall_ms,
               "api": "solvers.iterator(SVRG(plus=True)) -> next(state) -> solution(state)"}
        pass_rows = rows_per_gpu if sharded else (N // world if strong else N)      # rows one rank streams per pass
        algo_bytes = pass_rows * ld * 8
        extra = {"svrg": {"inner_steps": inner_steps, "us_per_inner_step": 1e3 * float(np.sum(seq_ms_list)) / inner_steps,
                          "inner_ms": [round(v, 3) for v in seq_ms_list], "pass_ms": [round(v, 3) for v in pass_ms_list],
                          "objective_start": f_start, "objective_end": f_end, "checksum_x": float(np.sum(np.abs(xs)))},
                 # the inner kernel is a dependency chain, not a bandwidth kernel: its "roofline" is the latency of one cluster-wide
                 # exchange per step (DSMEM store → remote mbarrier → wake-up ≈ 240 cycles of the ≈ 610-cycle step at 1.965 GHz)
                 "inner_kernel": {"bound": "latency", "kernel": "seq_kernel (persistent 8-CTA cluster, one launch per epoch)",
                                  "us_per_step": 1e3 * float(np.sum(seq_ms_list)) / inner_steps,
                                  "share_of_step_time": float(np.sum(seq_ms_list)) / ms,
                                  "exchange_floor_us": floor["lone_mode1_ns"] / 1e3,
                                  "exchange_floor": floor, "smids": smids,
                                  "us_per_step_by_epoch": [round(1e3 * v / m_of(k, N), 4) for k, v in enumerate(seq_ms_list)],
                                  "sm_mhz_by_epoch": [round(v, 1) for v in seq_mhz_list],
                                  "cycles_per_step_by_epoch": [round(1e3 * v / m_of(k, N) * f, 1) for k, (v, f) in enumerate(zip(seq_ms_list, seq_mhz_list))],
                                  "algorithmic_bytes_per_step": 8 * d + 32,
                                  "achieved_gbs": (8 * d + 32) * inner_steps / float(np.sum(seq_ms_list)) / 1e6},
                 "full_gradient": {"rows_per_gpu": pass_rows, "kernel_ms": pass_ms,
                                   "aggregate_gbs": (world if replicas else 1) * N * ld * 8 / pass_ms / 1e6,
                                   "gbs_per_gpu": algo_bytes / pass_ms / 1e6, "frac_of_8TBs": algo_bytes / pass_ms / 1e6 / 8000.0}}
        if replicas:
            # The part of the path that shards: the same full-gradient pass row-windowed over the ranks (N/G rows each) +
            # one NCCL allreduce of the d-vector, timed after the solves (SVRG_basic.jl:58-63 over a sharded F).
            obj = [np.asarray(xs, dtype=np.float64).tobytes() if rank == 0 else None]   # every replica ended somewhere else:
            dist.broadcast_object_list(obj, 0)                                          # the pass is evaluated at rank 0's x
            xs = np.frombuffer(obj[0], dtype=np.float64).copy()
            g_local = e.full_gradient(xs, 1.0 / N)                  # this rank alone, all N rows, no exchange
            init_comm()
            lo, hi = (rank * N) // world, ((rank + 1) * N) // world
            e.set_pass_window(lo, hi - lo)
            g_sharded = e.full_gradient(xs, 1.0 / N)                # N/G rows per rank + exchange
            rel_err = float(np.linalg.norm(g_sharded - g_local) / np.linalg.norm(g_local))
            bitwise = same_on_all_ranks(g_sharded)
            e.set_vec(L.VEC_X, xs)
            for _ in range(W):
                e.full_gradient(None, 1.0 / N, out=False)
            barrier()
            e.timer_begin()
            for _ in range(K):                                      # timed: back-to-back passes, no host synchronisation between them
                e.full_gradient(None, 1.0 / N, out=False)
            sh_ms = max_over_ranks(e.timer_end()) / K
            barrier()
            sh_ms_list, tail_ms_list = [], []
            for _ in range(K):                                      # per-kernel device times (reading them synchronises: separate loop)
                e.full_gradient(None, 1.0 / N, out=False)
                tm = e.last_timing()
                sh_ms_list.append(tm.last_pass_ms)
                tail_ms_list.append(tm.last_tail_ms)
            barrier()
            sh_kernel_ms = max_over_ranks(float(np.mean(sh_ms_list)))
            extra["full_gradient_sharded"] = {
                "what": f"one full-gradient pass over the C3 rows, row-windowed over {world} GPUs + exchange of the d-vector "
                        f"({'one-shot peer-memory exchange fused into the tail kernel of the pass' if os.environ.get('CIAO_BENCH_EXCHANGE', 'p2p') == 'p2p' else 'ncclAllReduce'})",
                "rows_per_gpu": hi - lo, "ms_per_pass": sh_ms, "kernel_ms": sh_kernel_ms, "passes": K,
                "tail_us": 1e3 * (sh_ms - sh_kernel_ms),
                "tail_kernel_us": 1e3 * max_over_ranks(float(np.mean(tail_ms_list))),
                "tail_what": "ms_per_pass - kernel_ms: launch gap + tail kernel (reduction of the CTA partials, exchange incl. the wait "
                             "for the slowest rank's pass, closing update); tail_kernel_us = that kernel alone, max over ranks",
                "rel_err_vs_local": max_over_ranks(rel_err), "rel_err_bound": 1e-13, "bitwise_equal_across_ranks": bool(bitwise),
                "aggregate_gbs": N * ld * 8 / sh_ms / 1e6, "kernel_gbs_per_gpu": (hi - lo) * ld * 8 / sh_kernel_ms / 1e6,
                "speedup_vs_local_pass": pass_ms / sh_ms}
        scaling = "strong" if args.workload == "svrgpp-strong" else "weak"   # default: per-GPU work fixed (one solve per GPU)
        if sharded:
            workload = (f"C4-style Lasso N={world}x2^{args.rows_log2} rows sharded over {world} GPUs, d={d} fp64 SVRG++ m=N/16*2^(k mod 5): "
                        f"sharded full-gradient pass + NCCL allreduce; inner epoch replicated, remote rows TMA-prefetched over NVLink (CUDA IPC)")
        else:
            workload = (f"C3 Lasso N=2^{args.rows_log2} d={d} fp64 SVRG++ gamma=1/(7 L_max) m=N/16*2^(k mod 5): persistent inner epoch + "
                        f"full-gradient pass"
                        + (f"; ONE solve on {world} GPUs: pass row-sharded + NCCL allreduce, inner epoch replicated" if strong else "")
                        + (f"; {world} independent solves, one per GPU (own index stream each; the sequential inner epoch does not shard), "
                           f"no data-path collective" if replicas else ""))

    clocks = sampler.summary() if rank == 0 else None
    achieved = algo_bytes / pass_ms / 1e6     # GB/s
    # dram__bytes_read.sum + dram__bytes_write.sum of row_pass_kernel from the committed `ncu --set full` capture
    # (profiles/ncu_row_pass_r1.csv: 137.90 GB read + 0.11 GB written per launch at N=2^22, d=4096, one GPU — the writes are
    # the 32 B/row step scalars a single-process pass leaves for the inner kernel; multi-rank passes write 12 MB); null otherwise
    traffic, traffic_src = None, None
    cap = os.path.join(ROOT, "profiles", "ncu_row_pass_traffic.json")
    if os.path.exists(cap) and rows_per_gpu == 1 << 22 and d == 4096 and (world == 1 or weak_pass or replicas):
        cj = json.load(open(cap))
        traffic = float(cj["dram_bytes_read"] + cj["dram_bytes_write"])
        traffic_src = cj["source"]
    line = {
        "metric": ("epochs/s (SAGA table-init passes, 2^22-row epochs)" if table_init else
                   "epochs/s (full-gradient passes, 2^22-row epochs)" if weak_pass else
                   "epochs/s (SVRG++ on row-sharded data, 2^22-row epochs)" if sharded else "epochs/s (Lasso 4M x 4096 fp64, SVRG++)"),
        "value": value, "unit": "epochs/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "l2": "inputs larger than L2 (137 GB of row records per pass; rows sampled at random)",
                   "epoch": "N component-gradient evaluations; step = (m + N)/N epochs", "seeds": [SEED_DATA, SEED_IDX]},
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "row_pass_kernel (SAGA table init: read rows + write table)" if table_init else "row_pass_kernel (full gradient)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes},
        "clocks": clocks, "setup": {"what": "ciao_gen_synthetic (rows generated in HBM)", "seconds": setup_s},
    }
    line.update(extra)
    e.close()
    if args.workload == "svrgpp" and not args.headline_only:
        if world == 1:
            line["c2"] = bench_c2(local, peak)
            line["c5"] = bench_c5(local, peak)
            line["setup_host_rows"] = bench_host_rows(local)
        else:
            line["c4"] = bench_c4(local, rank, world, peak, d, K, W, init_comm, barrier, max_over_ranks, same_on_all_ranks)
            try:
                line["c2_sharded_minibatch"] = bench_sharded_minibatch(local, rank, world, init_comm, barrier, max_over_ranks, same_on_all_ranks)
            except Exception as ex:      # an auxiliary figure must not take the headline line down with it (the error is the same on all ranks)
                line["c2_sharded_minibatch"] = {"error": str(ex)[:300]}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_reference(args, N, d, 5, 1)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

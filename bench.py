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
def cpu_reference(args, N_full, d, steps, warmup, seed_idx=SEED_IDX):
    """Times the oracle's SVRG++ outer iteration on a bounded sample (2^cpu_rows_log2 rows of the
    same generator, m = N_s/16·2^k) on one host core; extrapolates per-evaluation cost to N_full."""
    from oracle import oracle as orc
    Ns = 1 << args.cpu_rows_log2
    A, b = orc.gen_rows(orc.SYN_LASSO, d, SEED_DATA, 0, Ns)
    p = orc.Problem(orc.LOSS_LS, A, b, np.full(Ns, float(Ns))).set_reg(orc.REG_NORML1, lam=Ns / 100.0)
    gamma = 1 / (7 * Ns * p.max_row_sqnorm())
    st = orc.SVRGState(p, np.zeros(d), gamma, m=Ns // 16, plus=True)
    rng = np.random.default_rng(seed_idx)
    evals, t_inner, t_pass, t_total = 0, 0.0, 0.0, 0.0
    for k in range(-warmup, steps):
        m = (Ns // 16) << (max(k, 0) % SCHEDULE) if k >= 0 else Ns // 16
        idx = rng.integers(1, Ns + 1, size=m, dtype=np.int64)
        st.m = m
        t0 = time.perf_counter()
        st.inner(idx)
        t1 = time.perf_counter()
        st.av[:] = p.full_gradient(st.z_full, 1.0 / Ns)
        t2 = time.perf_counter()
        if k >= 0:
            evals += m + Ns
            t_inner += t1 - t0
            t_pass += t2 - t1
            t_total += t2 - t0
    per_eval = t_total / evals
    return {"value": 1.0 / (per_eval * N_full), "unit": "epochs/s", "cores": 1, "kind": "port",
            "sample": f"oracle (C restatement of the single-threaded reference; Julia is not installed) SVRG++ outer "
                      f"iterations on {Ns} rows x {d} (same generator), {steps} steps, {evals} component gradients in "
                      f"{t_total:.2f} s on 1 core (the reference has no threading; its BLAS calls are 1 x d gemv); "
                      f"extrapolated per component gradient to N = {N_full}",
            "us_per_inner_step": 1e6 * t_inner / max(1, evals - steps * Ns),
            "us_per_pass_row": 1e6 * t_pass / (steps * Ns), "seconds": t_total}, t_total / steps


def _reference_replica(job):
    args, N, d, steps, warmup, r = job
    return cpu_reference(args, N, d, steps, warmup, seed_idx=SEED_IDX + r)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, d = 1 << args.rows_log2, args.d
    steps = max(1, min(args.steps, 5))
    G = args.gpus if args.workload == "svrgpp" else 1
    if G > 1:
        # the GPU arm at N > 1 runs G independent solves, one per GPU: the CPU arm runs the same G solves concurrently,
        # one single-threaded reference loop per host core (the reference itself has no threading), 2 GiB of rows each
        import multiprocessing as mp
        args.cpu_rows_log2 = min(args.cpu_rows_log2, 16)
        with mp.get_context("spawn").Pool(G) as pool:
            res = pool.map(_reference_replica, [(args, N, d, steps, min(args.warmup, 1), r) for r in range(G)])
        s_per_step = max(r[1] for r in res)                    # the job ends with its slowest replica
        cb = dict(res[0][0])
        cb["value"] = G * min(r[0]["value"] for r in res)      # whole job: G solves at the pace of the slowest
        cb["cores"] = G
        cb["sample"] = f"{G} concurrent replicas, one per host core, each: " + cb["sample"]
    else:
        cb, s_per_step = cpu_reference(args, N, d, steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": "epochs/s (Lasso 4M x 4096 fp64, SVRG++)", "value": cb["value"], "unit": "epochs/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * s_per_step,
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
    def init_comm():
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(Engine.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        e.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

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
        pass_ms_list, seq_ms_list = [], []
        e.timer_begin()
        for k in range(K):
            e.svrg_epoch(idx_dev[k].data_ptr(), m_of(k, N))
            tm = e.last_timing()
            pass_ms_list.append(tm.last_pass_ms)
            seq_ms_list.append(tm.last_seq_ms)
        ms = max_over_ranks(e.timer_end())
        barrier()
        launches = e.last_timing().launches - l0
        epoch_rows = rows_per_gpu if sharded else N      # sharded: an "epoch" stays 2^22 component gradients
        epochs_total = sum((m_of(k, N) + N) / epoch_rows for k in range(K)) * (world if replicas else 1)
        value = epochs_total / (ms / 1e3)
        pass_ms = max_over_ranks(float(np.mean(pass_ms_list)))
        inner_steps = sum(m_of(k, N) for k in range(K))
        x_dev_path = e.get_vec(L.VEC_Z_FULL)
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
                                  "exchange_floor_us": 240 / 1965.0, "algorithmic_bytes_per_step": 8 * d + 32,
                                  "achieved_gbs": (8 * d + 32) * inner_steps / float(np.sum(seq_ms_list)) / 1e6},
                 "full_gradient": {"rows_per_gpu": pass_rows, "kernel_ms": pass_ms,
                                   "aggregate_gbs": (world if replicas else 1) * N * ld * 8 / pass_ms / 1e6,
                                   "gbs_per_gpu": algo_bytes / pass_ms / 1e6, "frac_of_8TBs": algo_bytes / pass_ms / 1e6 / 8000.0}}
        if replicas:
            # The part of the path that shards: the same full-gradient pass row-windowed over the ranks (N/G rows each) +
            # one NCCL allreduce of the d-vector, timed after the solves (SVRG_basic.jl:58-63 over a sharded F).
            init_comm()
            lo, hi = (rank * N) // world, ((rank + 1) * N) // world
            e.set_pass_window(lo, hi - lo)
            e.set_vec(L.VEC_X, xs)
            for _ in range(W):
                e.full_gradient(None, 1.0 / N, out=False)
            barrier()
            sh_ms_list = []
            e.timer_begin()
            for _ in range(K):
                e.full_gradient(None, 1.0 / N, out=False)
                sh_ms_list.append(e.last_timing().last_pass_ms)
            sh_ms = max_over_ranks(e.timer_end()) / K
            barrier()
            sh_kernel_ms = max_over_ranks(float(np.mean(sh_ms_list)))
            extra["full_gradient_sharded"] = {
                "what": f"one full-gradient pass over the C3 rows, row-windowed over {world} GPUs + NCCL allreduce of the d-vector",
                "rows_per_gpu": hi - lo, "ms_per_pass": sh_ms, "kernel_ms": sh_kernel_ms, "passes": K,
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
    traffic = ((138.011e9 if (world == 1 or replicas) else 137.899e9)
               if (rows_per_gpu == 1 << 22 and d == 4096 and (world == 1 or weak_pass or replicas)) else None)
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
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes},
        "clocks": clocks, "setup": {"what": "ciao_gen_synthetic (rows generated in HBM)", "seconds": setup_s},
    }
    line.update(extra)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_reference(args, N, d, 5, 1)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

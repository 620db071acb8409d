"""Thin object wrapper over one ``ciao_ctx`` (one GPU).  Every method is one C-ABI call."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import CiaoError, check, f64arr, i64arr, ptr  # noqa: F401


class Engine:
    def __init__(self, device: int = 0):
        self.lib = L.load()
        self.h = C.c_void_p()
        check(self.lib.ciao_create(C.byref(self.h), device))
        self.device = device
        self.N = self.d = self.n_rows = 0

    # -- lifetime -------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.ciao_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        check(self.lib.ciao_sync(self.h))

    # -- problem --------------------------------------------------------------
    def set_rows(self, loss_kind, A, b, scale=1.0, N_total=None, row0=0):
        A = f64arr(A)
        n_rows, d = A.shape
        b = f64arr(b)
        assert b.shape == (n_rows,)
        sv = None if np.isscalar(scale) else f64arr(scale)
        N_total = n_rows if N_total is None else N_total
        check(self.lib.ciao_set_rows(self.h, loss_kind, N_total, row0, n_rows, d, ptr(A), A.strides[0] // 8, ptr(b),
                                     ptr(sv), float(scale) if sv is None else 0.0))
        self.N, self.d, self.n_rows = N_total, d, n_rows

    def set_row_blocks(self, loss_kind, A, b, M, scale=1.0):
        """Components that are M×d blocks: A is (N·M)×d, b has N·M entries, scale N entries or a scalar."""
        A, b = f64arr(A), f64arr(b)
        rows, d = A.shape
        M = int(M)
        assert rows % M == 0 and b.shape == (rows,)
        N = rows // M
        sv = None if np.isscalar(scale) else f64arr(scale)
        check(self.lib.ciao_set_row_blocks(self.h, loss_kind, N, M, d, ptr(A), A.strides[0] // 8, ptr(b), ptr(sv),
                                           float(scale) if sv is None else 0.0))
        self.N, self.d, self.n_rows = N, d, N

    def set_blocks(self, Qdiag, qlin, box, eta):
        Q, q = f64arr(Qdiag), f64arr(qlin)
        N, n = Q.shape
        assert q.shape == (N, n)
        check(self.lib.ciao_set_blocks(self.h, N, n, ptr(Q), n, ptr(q), n, float(box[0]), float(box[1]), float(eta)))
        self.N, self.d, self.n_rows = N, n, N

    def set_reg(self, kind, *params):
        if kind == L.REG_INDBOX and len(params) == 2 and (np.ndim(params[0]) > 0 or np.ndim(params[1]) > 0):
            lo = np.broadcast_to(np.asarray(params[0], dtype=np.float64), (self.d,))
            hi = np.broadcast_to(np.asarray(params[1], dtype=np.float64), (self.d,))
            p = f64arr(np.concatenate([lo, hi]))
        else:
            p = f64arr(np.asarray(params, dtype=np.float64).reshape(-1))
        check(self.lib.ciao_set_reg(self.h, kind, ptr(p) if p.size else None, p.size))

    def set_row_interleave(self, block_rows, rank, world):
        """Interleaved row shards: this context holds the blocks of `block_rows` rows number rank, rank + world, … (before the rows are set)."""
        check(self.lib.ciao_set_row_interleave(self.h, block_rows, rank, world))

    def gen_synthetic(self, kind, N_total, d, seed, scale=1.0, row0=0, n_rows=None):
        n_rows = N_total if n_rows is None else n_rows
        check(self.lib.ciao_gen_synthetic(self.h, kind, N_total, row0, n_rows, d, seed, float(scale)))
        self.N, self.d, self.n_rows = N_total, d, n_rows

    @staticmethod
    def gen_host(kind, d, seed, row0, n_rows):
        A = np.empty((n_rows, d))
        rhs = np.empty(n_rows)
        check(L.load().ciao_gen_host(kind, d, seed, row0, n_rows, ptr(A), ptr(rhs)))
        return A, (None if kind == L.SYNTH_SHARING else rhs)

    # -- multi-GPU -------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(L.load().ciao_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = C.create_string_buffer(uid, 128)
        check(self.lib.ciao_comm_init(self.h, buf, rank, world))

    def comm_p2p_handle(self) -> bytes:
        """128-byte blob of this context's exchange arena; all-gather the blobs in rank order, then comm_p2p_attach."""
        buf = C.create_string_buffer(128)
        check(self.lib.ciao_comm_p2p_handle(self.h, buf))
        return buf.raw

    def comm_p2p_attach(self, rank: int, world: int, handles):
        blob = C.create_string_buffer(b"".join(handles), 128 * len(handles))
        check(self.lib.ciao_comm_p2p_attach(self.h, int(rank), int(world), blob))

    def rows_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(self.lib.ciao_rows_ipc_handle(self.h, buf))
        return buf.raw

    def attach_peer_rows(self, handles, row0, n_rows, my_shard):
        """handles: list of 64-byte IPC handles in shard order (the entry of my_shard is ignored)."""
        blob = C.create_string_buffer(b"".join(handles), 64 * len(handles))
        r0, nr = i64arr(row0), i64arr(n_rows)
        check(self.lib.ciao_attach_peer_rows(self.h, len(handles), blob, ptr(r0), ptr(nr), int(my_shard)))

    def set_pass_window(self, row_lo, n):
        check(self.lib.ciao_set_pass_window(self.h, int(row_lo), int(n)))

    # -- passes ----------------------------------------------------------------
    def full_gradient(self, x=None, scale=1.0, out=True):
        xv = None if x is None else f64arr(x)
        o = np.empty(self.d) if out else None
        check(self.lib.ciao_full_gradient(self.h, ptr(xv), float(scale), ptr(o)))
        return o

    def objective(self, x):
        xv = f64arr(x)
        f, g = C.c_double(), C.c_double()
        check(self.lib.ciao_objective(self.h, ptr(xv), C.byref(f), C.byref(g)))
        return f.value, g.value

    def max_row_sqnorm(self):
        o = C.c_double()
        check(self.lib.ciao_max_row_sqnorm(self.h, C.byref(o)))
        return o.value

    # -- solvers ---------------------------------------------------------------
    def svrg_init(self, x0, gamma, plus=False):
        check(self.lib.ciao_svrg_init(self.h, ptr(f64arr(x0)), float(gamma), int(plus)))

    def svrg_epoch(self, idx, m=None):
        """idx: int64 numpy array (host), an int device address, or None (staged indices)."""
        if isinstance(idx, np.ndarray):
            idx = i64arr(idx)
            m = len(idx) if m is None else m
        check(self.lib.ciao_svrg_epoch(self.h, ptr(idx), int(m)))

    def saga_init(self, x0, gamma, sag=False):
        check(self.lib.ciao_saga_init(self.h, ptr(f64arr(x0)), float(gamma), int(sag)))

    def saga_steps(self, idx, K=None):
        if isinstance(idx, np.ndarray):
            idx = i64arr(idx)
            K = len(idx) if K is None else K
        check(self.lib.ciao_saga_steps(self.h, ptr(idx), int(K)))

    def finito_init(self, x0, gamma_N, hat_gamma):
        check(self.lib.ciao_finito_init(self.h, ptr(f64arr(x0)), ptr(f64arr(gamma_N)), float(hat_gamma)))

    def finito_steps(self, idx, batch_ptr):
        idx, bp = (i64arr(idx) if isinstance(idx, np.ndarray) else idx), i64arr(batch_ptr)
        check(self.lib.ciao_finito_steps(self.h, ptr(idx), ptr(bp), len(bp) - 1))

    def lfinito_init(self, x0, gamma_N, hat_gamma):
        check(self.lib.ciao_lfinito_init(self.h, ptr(f64arr(x0)), ptr(f64arr(gamma_N)), float(hat_gamma)))

    def lfinito_outer(self, batch_order, r):
        o = i64arr(batch_order)
        check(self.lib.ciao_lfinito_outer(self.h, ptr(o), len(o), int(r)))

    def finito_adaptive_init(self, x0, alpha=0.999, tol_b=1e-9, perturb=None):
        """perturb(i, t) -> d-vector `rand(t * [-1, 1], size(x0))` from the host's RNG: the random restart of the stepsize
        estimate for components with ∇f_i(x0 + 1) == ∇f_i(x0) (Finito_adaptive.jl:77-83); None → such a problem is refused."""
        x0 = f64arr(x0)
        if perturb is None:
            check(self.lib.ciao_finito_adaptive_init(self.h, ptr(x0), float(alpha), float(tol_b)))
            return
        d = self.d

        def _cb(_user, i1, t, out):
            try:
                np.ctypeslib.as_array(out, shape=(d,))[:] = x0 + np.asarray(perturb(int(i1), int(t)), dtype=np.float64)
                return 0
            except Exception:
                return 1

        cb = L.PERTURB_FN(_cb)
        check(self.lib.ciao_finito_adaptive_init_cb(self.h, ptr(x0), float(alpha), float(tol_b), C.cast(cb, C.c_void_p), None))

    def finito_adaptive_steps(self, idx, K=None) -> int:
        """Returns the number of steps completed (< K ⇔ `return nothing`, Finito_adaptive.jl:124-127)."""
        if isinstance(idx, np.ndarray):
            idx = i64arr(idx)
            K = len(idx) if K is None else K
        done = C.c_int64()
        check(self.lib.ciao_finito_adaptive_steps(self.h, ptr(idx), int(K), C.byref(done)))
        return done.value

    def finito_adaptive_get(self, gamma=True, fi_x=False, coef=False):
        """(γ, f_i(x_i), c_i, γ̂, number of γ reductions); arrays not asked for are None."""
        outs = [np.empty(self.N) if w else None for w in (gamma, fi_x, coef)]
        hg, nbt = C.c_double(), C.c_int64()
        check(self.lib.ciao_finito_adaptive_get(self.h, ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), C.byref(hg), C.byref(nbt)))
        return outs[0], outs[1], outs[2], hg.value, nbt.value

    def proshi_init(self, x0, gamma_N, hat_gamma):
        check(self.lib.ciao_proshi_init(self.h, ptr(f64arr(x0)), ptr(f64arr(gamma_N)), float(hat_gamma)))

    def proshi_steps(self, idx, batch_ptr):
        idx, bp = (i64arr(idx) if isinstance(idx, np.ndarray) else idx), i64arr(batch_ptr)
        check(self.lib.ciao_proshi_steps(self.h, ptr(idx), ptr(bp), len(bp) - 1))

    def proshi_solution(self, out=None):
        check(self.lib.ciao_proshi_solution(self.h, ptr(out)))
        return out

    # -- state -------------------------------------------------------------------
    def get_vec(self, which, out=None):
        out = np.empty(self.d) if out is None else out
        check(self.lib.ciao_get_vec(self.h, which, ptr(out), self.d))
        return out

    def set_vec(self, which, x):
        check(self.lib.ciao_set_vec(self.h, which, ptr(f64arr(x)), self.d))

    def get_table_rows(self, i0=0, n=None, out=None):
        n = self.n_rows - i0 if n is None else n
        out = np.empty((n, self.d)) if out is None else out
        check(self.lib.ciao_get_table_rows(self.h, i0, n, ptr(out)))
        return out

    def set_table_rows(self, rows, i0=0):
        rows = f64arr(rows)
        assert rows.ndim == 2 and rows.shape[1] == self.d
        check(self.lib.ciao_set_table_rows(self.h, int(i0), rows.shape[0], ptr(rows)))

    def solver_restore(self, algo, gamma=0.0, flag=False, gamma_N=None, hat_gamma=0.0):
        """algo: 1 SVRG, 2 SAGA, 3 Finito, 4 LFinito, 5 ProShI — state without the init pass; then set_vec / set_table_rows."""
        g = None if gamma_N is None else f64arr(gamma_N)
        check(self.lib.ciao_solver_restore(self.h, int(algo), float(gamma), int(flag), ptr(g), float(hat_gamma)))

    def table_colsum(self):
        o = np.empty(self.d)
        check(self.lib.ciao_table_colsum(self.h, ptr(o)))
        return o

    # -- measurement ---------------------------------------------------------------
    def stage_indices(self, idx):
        idx = i64arr(idx)
        check(self.lib.ciao_stage_indices(self.h, ptr(idx), len(idx)))

    def timer_begin(self):
        check(self.lib.ciao_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        check(self.lib.ciao_timer_end(self.h, C.byref(ms)))
        return ms.value

    def last_timing(self) -> L.Timing:
        t = L.Timing()
        check(self.lib.ciao_last_timing(self.h, C.byref(t)))
        return t

    def last_seq_placement(self):
        """SM ids of the CTAs of the last sequential cluster kernel."""
        sm, n = (C.c_int * 16)(), C.c_int()
        check(self.lib.ciao_last_seq_placement(self.h, sm, C.byref(n)))
        return list(sm[: n.value])

    def last_seq_clock_mhz(self) -> float:
        """SM clock (MHz) the last sequential cluster kernel actually ran at: its own cycle counter over the global timer."""
        cyc, ns = C.c_int64(), C.c_int64()
        check(self.lib.ciao_last_seq_clock(self.h, C.byref(cyc), C.byref(ns)))
        return 1e3 * cyc.value / ns.value if ns.value > 0 else 0.0

    def measure_exchange(self, cluster=8, warps=4, iters=100000, mode=1, max_clusters=74):
        """Per-cluster (ns per round, SM cycles per round, SM ids) of the cluster-exchange floor (seq_floor.cu)."""
        ns, cyc = (C.c_float * max_clusters)(), (C.c_float * max_clusters)()
        sm, n = (C.c_int * (max_clusters * cluster))(), C.c_int()
        check(self.lib.ciao_measure_exchange(self.h, cluster, warps, iters, mode, max_clusters, C.byref(n), ns, cyc, sm))
        return [(ns[k], cyc[k], list(sm[k * cluster:(k + 1) * cluster])) for k in range(n.value)]

    def set_tuning(self, pass_threads=0, pass_stages=0, pass_ctas_per_sm=0, seq_cluster=0, seq_threads=0):
        check(self.lib.ciao_set_tuning(self.h, pass_threads, pass_stages, pass_ctas_per_sm, seq_cluster, seq_threads))

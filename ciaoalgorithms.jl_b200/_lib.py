"""ctypes binding of libciao_cuda.so — exactly the symbols include/ciao_cuda.h declares.

The Julia shim binds the same symbols with ``ccall``.  There is
no CPU fallback: if the library is missing ``load()`` raises, and every compute
call raises ``CiaoError`` when no GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libciao_cuda.so")

OK = 0
ERR_NAMES = {-1: "CIAO_ERR_INVALID", -2: "CIAO_ERR_CUDA", -3: "CIAO_ERR_STATE",
             -4: "CIAO_ERR_UNSUPPORTED", -5: "CIAO_ERR_COMM", -6: "CIAO_ERR_OOM"}

LOSS_LS, LOSS_LOGISTIC, LOSS_DIAGQUAD = 0, 1, 2
REG_ZERO, REG_NORML1, REG_INDBOX, REG_NORML1_PAIRS = 0, 1, 2, 3
VEC_Z, VEC_Z_FULL, VEC_W, VEC_AV, VEC_X = 0, 1, 2, 3, 4
SYNTH_LASSO, SYNTH_LOGISTIC, SYNTH_SHARING = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_ctx = C.c_void_p
i64, f64, i32 = C.c_int64, C.c_double, C.c_int


class Timing(C.Structure):
    _fields_ = [("last_pass_ms", C.c_float), ("last_seq_ms", C.c_float),
                ("last_pass_bytes", C.c_int64), ("last_seq_steps", C.c_int64), ("launches", C.c_int64),
                ("last_tail_ms", C.c_float)]


# name → (restype, argtypes); must list every symbol of include/ciao_cuda.h
SIGNATURES = {
    "ciao_version": (i32, []),
    "ciao_last_error": (C.c_char_p, []),
    "ciao_device_count": (i32, [C.POINTER(i32)]),
    "ciao_create": (i32, [C.POINTER(_ctx), i32]),
    "ciao_destroy": (i32, [_ctx]),
    "ciao_sync": (i32, [_ctx]),
    "ciao_set_rows": (i32, [_ctx, i32, i64, i64, i64, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p, f64]),
    "ciao_set_row_blocks": (i32, [_ctx, i32, i64, i64, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p, f64]),
    "ciao_set_row_interleave": (i32, [_ctx, i64, i32, i32]),
    "ciao_set_blocks": (i32, [_ctx, i64, i64, C.c_void_p, i64, C.c_void_p, i64, f64, f64, f64]),
    "ciao_set_reg": (i32, [_ctx, i32, C.c_void_p, i64]),
    "ciao_gen_synthetic": (i32, [_ctx, i32, i64, i64, i64, i64, C.c_uint64, f64]),
    "ciao_gen_host": (i32, [i32, i64, C.c_uint64, i64, i64, C.c_void_p, C.c_void_p]),
    "ciao_comm_unique_id": (i32, [C.c_void_p]),
    "ciao_comm_init": (i32, [_ctx, C.c_void_p, i32, i32]),
    "ciao_comm_p2p_handle": (i32, [_ctx, C.c_void_p]),
    "ciao_comm_p2p_attach": (i32, [_ctx, i32, i32, C.c_void_p]),
    "ciao_set_pass_window": (i32, [_ctx, i64, i64]),
    "ciao_rows_ipc_handle": (i32, [_ctx, C.c_void_p]),
    "ciao_attach_peer_rows": (i32, [_ctx, i32, C.c_void_p, C.c_void_p, C.c_void_p, i32]),
    "ciao_full_gradient": (i32, [_ctx, C.c_void_p, f64, C.c_void_p]),
    "ciao_objective": (i32, [_ctx, C.c_void_p, _dp, _dp]),
    "ciao_max_row_sqnorm": (i32, [_ctx, _dp]),
    "ciao_svrg_init": (i32, [_ctx, C.c_void_p, f64, i32]),
    "ciao_svrg_epoch": (i32, [_ctx, C.c_void_p, i64]),
    "ciao_saga_init": (i32, [_ctx, C.c_void_p, f64, i32]),
    "ciao_saga_steps": (i32, [_ctx, C.c_void_p, i64]),
    "ciao_finito_init": (i32, [_ctx, C.c_void_p, C.c_void_p, f64]),
    "ciao_finito_steps": (i32, [_ctx, C.c_void_p, C.c_void_p, i64]),
    "ciao_lfinito_init": (i32, [_ctx, C.c_void_p, C.c_void_p, f64]),
    "ciao_lfinito_outer": (i32, [_ctx, C.c_void_p, i64, i64]),
    "ciao_finito_adaptive_init": (i32, [_ctx, C.c_void_p, f64, f64]),
    "ciao_finito_adaptive_init_cb": (i32, [_ctx, C.c_void_p, f64, f64, C.c_void_p, C.c_void_p]),
    "ciao_finito_adaptive_steps": (i32, [_ctx, C.c_void_p, i64, _ip]),
    "ciao_finito_adaptive_get": (i32, [_ctx, C.c_void_p, C.c_void_p, C.c_void_p, _dp, _ip]),
    "ciao_proshi_init": (i32, [_ctx, C.c_void_p, C.c_void_p, f64]),
    "ciao_proshi_steps": (i32, [_ctx, C.c_void_p, C.c_void_p, i64]),
    "ciao_proshi_solution": (i32, [_ctx, C.c_void_p]),
    "ciao_get_vec": (i32, [_ctx, i32, C.c_void_p, i64]),
    "ciao_set_vec": (i32, [_ctx, i32, C.c_void_p, i64]),
    "ciao_get_table_rows": (i32, [_ctx, i64, i64, C.c_void_p]),
    "ciao_set_table_rows": (i32, [_ctx, i64, i64, C.c_void_p]),
    "ciao_solver_restore": (i32, [_ctx, i32, f64, i32, C.c_void_p, f64]),
    "ciao_table_colsum": (i32, [_ctx, C.c_void_p]),
    "ciao_jlrng_next_u64": (i32, [C.c_void_p, C.c_void_p, i64]),
    "ciao_jlrng_rand_range": (i32, [C.c_void_p, i64, C.c_void_p, i64]),
    "ciao_jlrng_randperm": (i32, [C.c_void_p, i64, C.c_void_p]),
    "ciao_jlrng_sample_norep": (i32, [C.c_void_p, i64, i64, C.c_void_p]),
    "ciao_stage_indices": (i32, [_ctx, C.c_void_p, i64]),
    "ciao_timer_begin": (i32, [_ctx]),
    "ciao_timer_end": (i32, [_ctx, C.POINTER(C.c_float)]),
    "ciao_last_timing": (i32, [_ctx, C.POINTER(Timing)]),
    "ciao_last_seq_placement": (i32, [_ctx, C.POINTER(i32), C.POINTER(i32)]),
    "ciao_last_seq_clock": (i32, [_ctx, _ip, _ip]),
    "ciao_measure_exchange": (i32, [_ctx, i32, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.POINTER(i32)]),
    "ciao_set_tuning": (i32, [_ctx, i32, i32, i32, i32, i32]),
}


PERTURB_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_double))   # ciao_perturb_fn


class CiaoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load():
    """dlopen libciao_cuda.so and declare every signature.  Raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise FileNotFoundError(
                f"{SO_PATH} not found — build it with `python ciaoalgorithms.jl_b200/build.py` "
                "(there is no CPU fallback)")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(code):
    if code != OK:
        raise CiaoError(code, load().ciao_last_error().decode("utf-8", "replace"))


def f64arr(a):
    """fp64 view/copy of a host array.  Complex-typed input is accepted when every imaginary part is zero — what the reference's
    complex test problems are (test_lasso.jl:3, :19: `C = rand(R, N, n)` is real, only the element type is complex), and for such
    data the reference's complex arithmetic keeps every imaginary part at exactly 0, so computing on the real parts gives the same
    bits.  Genuinely complex data never comes through here: operators.pack_F realifies it first (realify_rows / realify_vec)."""
    a = np.asarray(a)
    if np.iscomplexobj(a):
        if np.any(a.imag != 0):
            raise TypeError("complex data with non-zero imaginary parts must be realified first (realify_rows / realify_vec)")
        a = a.real
    return np.ascontiguousarray(a, dtype=np.float64)


def realify_vec(z):
    """complex d-vector → real 2d-vector (re_0, im_0, re_1, im_1, …): the memory layout of a Julia Vector{ComplexF64}."""
    z = np.asarray(z, dtype=np.complex128).reshape(-1)
    return np.ascontiguousarray(np.column_stack([z.real, z.imag]).reshape(-1))


def complexify_vec(x):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    return x[0::2] + 1j * x[1::2]


def realify_rows(A):
    """complex m×d matrix → real 2m×2d block [Re; Im] acting on realify_vec(x): rows 2r, 2r+1 give Re and Im of (A x)_r, and
    its transpose applied to (Re res, Im res) gives realify_vec(Aᴴ res) — the LeastSquares gradient of the complex problem."""
    A = np.atleast_2d(np.asarray(A, dtype=np.complex128))
    m, d = A.shape
    out = np.empty((2 * m, 2 * d))
    out[0::2, 0::2], out[0::2, 1::2] = A.real, -A.imag
    out[1::2, 0::2], out[1::2, 1::2] = A.imag, A.real
    return out


def i64arr(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def ptr(a):
    """void* of a numpy array, an int device/host address, or None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)

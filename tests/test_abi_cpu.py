"""CPU-side checks of the boundary: the C-ABI library loads without a GPU, exports every
symbol include/ciao_cuda.h declares (and the ctypes table lists exactly those), refuses to
compute without a device, and the host generator agrees with the oracle's."""
import ctypes
import os
import re

import numpy as np
import pytest

from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200 import build as B
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ciao_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ciao_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    so = B.build()
    lib = ctypes.CDLL(so)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in ciao_cuda.h but not exported"
    assert sorted(L.SIGNATURES) == syms, "ctypes table and header disagree"
    assert L.load().ciao_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ciaoalgorithms_jl_b200.engine import CiaoError, Engine
    with pytest.raises(CiaoError) as ei:
        Engine(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_host_generator_matches_oracle_generator():
    from ciaoalgorithms_jl_b200.engine import Engine
    for kind, okind in [(L.SYNTH_LASSO, orc.SYN_LASSO), (L.SYNTH_LOGISTIC, orc.SYN_LOGISTIC), (L.SYNTH_SHARING, orc.SYN_SHARING)]:
        A, r = Engine.gen_host(kind, 40, 123, 5, 17)
        A2, r2 = orc.gen_rows(okind, 40, 123, 5, 17)
        assert np.array_equal(A, A2)
        assert (r is None and r2 is None) or np.array_equal(r, r2)


def test_sass_shows_tma_and_cluster_instructions():
    """The streaming pass and the sequential kernels really use TMA bulk copies (UBLKCP),
    mbarriers (SYNCS) and cluster barriers — B200_PROFILING.md 'What proves a Blackwell-native kernel'."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", B.build()], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass and "UCGABAR" in sass
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "EF_CUDA_SM100" in sass


def test_ctypes_signatures_match_the_header_prototypes():
    """Arity, scalar widths and pointer-ness of every entry of the ctypes table against include/ciao_cuda.h — the same
    check test_julia_shim_static.py makes for the Julia shim's ccalls."""
    from test_julia_shim_static import header_prototypes
    C = ctypes
    scalars = {"int": C.c_int, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "double": C.c_double, "float": C.c_float}
    rets = {"int": C.c_int, "const char *": C.c_char_p}
    protos = header_prototypes()
    assert sorted(protos) == sorted(L.SIGNATURES)
    for name, (res, args) in L.SIGNATURES.items():
        c_ret, c_args = protos[name]
        assert rets[c_ret] is res, f"{name}: return type {c_ret}"
        assert len(args) == len(c_args), f"{name}: {len(c_args)} parameters in the header, {len(args)} in the ctypes table"
        for k, (ct, at) in enumerate(zip(c_args, args)):
            if ct.endswith("*") or ct == "ciao_perturb_fn":      # data pointers and the one function-pointer type
                is_ptr = at in (C.c_void_p, C.c_char_p) or hasattr(at, "contents") or issubclass(at, C._Pointer)
                assert is_ptr, f"{name}: argument {k + 1} is `{ct}` in the header but {at} in the ctypes table"
            else:
                assert scalars[ct] is at, f"{name}: argument {k + 1} is `{ct}` in the header but {at} in the ctypes table"

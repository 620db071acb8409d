"""Builds libciao_cuda.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python ciaoalgorithms.jl_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libciao_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--fmad=true",
    "-Xptxas", "-v",
]
# translation units (abi.cu is a unity file that includes pass/gen/indices/proshi/comm)
UNITS = ["abi.cu", "seq_svrg.cu", "seq_saga.cu", "seq_finito.cu", "seq_lfinito.cu", "seq_adaptive.cu", "seq_floor.cu", "jlrng.cu"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))
                  + glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(f) > t for f in sources())


def build(force: bool = False, verbose: bool = False, defines=(), so: str = SO) -> str:
    if not (force or stale() or not os.path.exists(so)):
        return so
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "..", "build", os.path.basename(so))   # repo-root build/ (git- and gpurun-ignored)
    os.makedirs(objdir, exist_ok=True)
    dflags = [f"-D{d}" for d in defines]

    def compile_unit(u):
        obj = os.path.join(objdir, u.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + dflags + ["-c", "-o", obj, os.path.join(CSRC, u)]
        return u, obj, cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(len(UNITS)) as ex:
        results = list(ex.map(compile_unit, UNITS))
    log = os.path.join(objdir, "build.log")
    with open(log, "w") as f:
        for u, obj, cmd, res in results:
            f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr + "\n")
    for u, obj, cmd, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {u} (exit {res.returncode}); see {log}")
    link = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", so] + \
           [r[1] for r in results] + ["-ldl"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc link failed")
    if verbose:
        print(open(log).read())
    return so


if __name__ == "__main__":
    if "--profile" in sys.argv:   # per-phase cycle counters in the sequential kernel (scripts/prof_seq.py)
        print(build(force=True, defines=("CIAO_SEQ_PROFILE",), so=os.path.join(HERE, "libciao_cuda_prof.so")))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

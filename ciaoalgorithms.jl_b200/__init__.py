"""B200-native engine for the iteration loops of CIAOAlgorithms.jl.

Host-side mirror of the reference's solver API (SVRG/SVRG++, SAGA/SAG,
Finito/MISO/DIAG incl. LFinito, ProShI) over the C-ABI library
``libciao_cuda.so`` (include/ciao_cuda.h), whose hot path is hand-written
sm_100a CUDA.  There is no CPU fallback: every compute entry point raises if
the CUDA library is missing or no GPU is present.
"""
from . import sampling  # noqa: F401  (pure host logic, importable without the CUDA library)

__all__ = ["sampling"]

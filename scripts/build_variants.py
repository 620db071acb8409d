"""Builds experiment variants of libciao_cuda (compile-time switches of seq_impl.cuh) next to the product library:
    python scripts/build_variants.py  →  ciaoalgorithms.jl_b200/libciao_cuda_<tag>.so"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import build as b
VARIANTS = {"e1": ("CIAO_SEQ_E1",), "spin": ("CIAO_SEQ_SPIN",)}
for tag in (sys.argv[1:] or VARIANTS):
    print(b.build(force=True, defines=VARIANTS[tag], so=os.path.join(b.HERE, f"libciao_cuda_{tag}.so")), flush=True)

"""Loader for the product package.

The package directory is named ``ciaoalgorithms.jl_b200`` (after the reference
repo); the dot makes it un-importable by a plain ``import`` statement, so it is
registered under the module name ``ciaoalgorithms_jl_b200``:

    import ciao_pkg; ciao = ciao_pkg.load()
    from ciaoalgorithms_jl_b200 import solvers        # works after load()
"""
import importlib.util
import os
import sys

NAME = "ciaoalgorithms_jl_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "ciaoalgorithms.jl_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(
        NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod

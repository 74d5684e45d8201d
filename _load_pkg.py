"""Registers the directory ``gpu-spmv_b200/`` (hyphenated, so not importable by
name) as the Python package ``gpu_spmv_b200``.  Used by tests/, bench.py and
__graft_entry__.py:

    from _load_pkg import load_pkg
    sp = load_pkg()
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "gpu-spmv_b200")
PKG_NAME = "gpu_spmv_b200"


def load_pkg():
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(
        PKG_NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(PKG_NAME, None)
        raise
    return mod

"""Locating and loading the native libraries.

``libqz_b200.so``   CUDA kernels + C ABI (built by ``__graft_entry__.build()`` with nvcc, sm_100a)
``libqz_harness.so`` C++ host library + scene harness, linked against ``libqz_b200.so``

There is no fallback: if the libraries are missing, or there is no CUDA device, loading or
the first compute call fails with an error that says so.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import os

PKG_DIR = Path(__file__).resolve().parent
# QZ_LIB_DIR selects another in-tree build of the same sources (tuning variants, profiles/)
LIB_DIR = Path(os.environ["QZ_LIB_DIR"]).resolve() if os.environ.get("QZ_LIB_DIR") else PKG_DIR / "_lib"


def native_paths() -> dict:
    return {"cuda": LIB_DIR / "libqz_b200.so", "harness": LIB_DIR / "libqz_harness.so"}


def load_cuda_library() -> ctypes.CDLL:
    path = native_paths()["cuda"]
    if not path.exists():
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); "
                           "there is no CPU fallback")
    return ctypes.CDLL(str(path))


def load_harness():
    """The product harness (host library over the CUDA C ABI)."""
    from .harness import Harness

    load_cuda_library()
    path = native_paths()["harness"]
    if not path.exists():
        raise RuntimeError(f"{path} is missing: run __graft_entry__.build(); there is no CPU fallback")
    return Harness(path, "qzh_")

"""polympc_b200 — B200-native batched SQP engine for PolyMPC's collocated-NLP hot path.

The product is ``libpolympc_b200.so`` (hand-written sm_100a kernels behind the C ABI of ``include/polympc_b200.h``);
this package only locates / builds / loads it and offers a numpy-level binding for tests and the benchmark.  There is
no CPU execution path: ``load()`` raises when the library is missing and every compute entry point of the library
returns ``PMB_ERR_NO_DEVICE`` when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

from .capi import CApi, PmbError  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpolympc_b200.so")
_api = None


def build(jobs: int | None = None) -> str:
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    jobs = jobs or os.cpu_count() or 4
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), f"-j{jobs}"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def load() -> CApi:
    """Load the CUDA library.  Fails loudly when it has not been built: there is no fallback."""
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise PmbError(f"{LIB_PATH} is missing: run `make -C polympc_b200/csrc` (or __graft_entry__.build()); "
                           "polympc_b200 has no CPU fallback")
        _api = CApi(ctypes.CDLL(LIB_PATH), "pmb_")
    return _api

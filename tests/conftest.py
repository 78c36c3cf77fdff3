"""Shared fixtures.  `-m "not gpu"`: oracle vs the reference's fixtures, warp-emulator build of the kernels vs oracle, ABI
and host logic.  `-m gpu`: the CUDA library through its C ABI vs the oracle (bit-exact) on a real B200."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_plugins = {}


def _load_dropin_plugin(kind):
    """tests/dropin: reference-style problem classes built as a run-time plug-in (a user translation unit) that registers
    `dropin_robot_5x3` / `dropin_cstr_5x2` with the already loaded engine library when it is dlopen'ed"""
    if kind in _plugins:
        return
    d = os.path.join(ROOT, "tests", "dropin")
    lib = os.path.join(d, "_build", f"libdropin_{kind}.so")
    if kind == "emu" or not os.path.exists(lib):
        subprocess.run(["make", "-C", d, kind], check=True, stdout=subprocess.DEVNULL)
    _plugins[kind] = ctypes.CDLL(lib)      # resolves the engine through its own DT_NEEDED entry: the already loaded library


@pytest.fixture(scope="session")
def orc():
    """CPU oracle (test infrastructure), prefix orc_."""
    from oracle import pyoracle
    return pyoracle.load()


@pytest.fixture(scope="session")
def emu():
    """The product's kernel sources compiled by g++ against the lock-step warp emulator (test infrastructure), prefix emu_."""
    from polympc_b200.capi import CApi
    d = os.path.join(ROOT, "tests", "warp_emu")
    lib = os.path.join(d, "_build", "libpmb_emu.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", d, f"-j{os.cpu_count() or 4}"], check=True, stdout=subprocess.DEVNULL)
    api = CApi(ctypes.CDLL(lib), "emu_")
    _load_dropin_plugin("emu")
    return api


@pytest.fixture(scope="session")
def pmb():
    """The product: libpolympc_b200.so (sm_100a).  Loading never falls back to anything else."""
    import polympc_b200
    api = polympc_b200.load()
    if api.device_count() > 0:
        _load_dropin_plugin("gpu")
    return api


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "casadi_robot_5x2.npz"))


def rel_inf(a, b):
    """SURVEY.md §8d iterate parity: max|a-b| / max(1, max|b|)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))

"""oracle/pyoracle.py — TEST INFRASTRUCTURE.  Loads the CPU oracle (prefix ``orc_``) through the same ctypes wrapper as the
product.  May only be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from polympc_b200.capi import CApi  # noqa: E402

LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libcasadi_robot.so")


def build(force: bool = False) -> None:
    """Compile the oracle (and, when /root/reference is present, the reference's CasADi C fixtures)."""
    if force or not os.path.exists(LIB_PATH) or os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)


_api = None


def load() -> CApi:
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        _api = CApi(lib, "orc_")
        lib.orc_set_num_threads.argtypes = [ctypes.c_int]
        lib.orc_get_num_threads.restype = ctypes.c_int
    return _api


NATIVE_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_native.so")
_native = None


def load_native() -> CApi:
    """the same oracle sources built `-O3 -march=native -ffp-contract=fast` against libm: a faster CPU baseline and the
    rounding-sensitivity control.  NOT bit-reproducible (it is tied to the host CPU); never the parity reference."""
    global _native
    if _native is None:
        if not os.path.exists(NATIVE_LIB_PATH):
            build(force=True)
        lib = ctypes.CDLL(NATIVE_LIB_PATH)
        _native = CApi(lib, "orc_")
        lib.orc_set_num_threads.argtypes = [ctypes.c_int]
    return _native


def set_num_threads(n: int) -> None:
    load().lib.orc_set_num_threads(int(n))
    if _native is not None:
        _native.lib.orc_set_num_threads(int(n))

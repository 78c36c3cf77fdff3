"""The C-ABI boundary: header <-> binding <-> built library (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from polympc_b200.capi import ABI_FUNCTIONS, CApi, Dims, PmbError, QpSettings, SqpSettings

HEADER = os.path.join(ROOT, "include", "polympc_b200.h")


def header_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pmb_[a-z0-9_]+)\s*\(", txt)))


def test_binding_lists_every_header_function():
    assert sorted("pmb_" + f for f in ABI_FUNCTIONS) == header_functions()


def test_header_cites_reference_interfaces():
    txt = open(HEADER).read()
    for cite in ("continuous_ocp.hpp", "sqp_base.hpp", "qp_base.hpp", "box_admm.hpp", "bfgs.hpp", "ebyshev.hpp", "mpc_wrapper.hpp"):
        assert cite in txt


def test_product_library_exports_every_symbol(pmb):
    for f in header_functions():
        assert hasattr(pmb.lib, f), f
    assert "sm_100a" in pmb.version()


def test_oracle_and_emulator_export_the_same_abi(orc, emu):
    for api in (orc, emu):
        for f in ABI_FUNCTIONS:
            assert hasattr(api.lib, api.prefix + f), api.prefix + f


def test_library_contains_sm100a_code():
    import subprocess
    import polympc_b200
    out = subprocess.run(["cuobjdump", "-lelf", polympc_b200.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_registry_and_dims(pmb, orc):
    names = pmb.problems()
    # dropin_*: reference-style classes compiled through include/polympc_compat/; their oracle twins are the hand-restated models
    assert [n for n in names if not n.startswith("dropin_")] == orc.problems()
    assert {"mobile_robot_6x2", "cstr_5x2", "kite_12x1", "robot_obstacle_5x2"} <= set(names)
    d = pmb.dims("mobile_robot_6x2")
    assert (d["NX"], d["NU"], d["NN"], d["N"], d["M"], d["DUAL"]) == (3, 2, 13, 65, 39, 104)
    d = pmb.dims("cstr_5x2")
    assert (d["N"], d["M"]) == (66, 44)
    d = pmb.dims("kite_12x1")
    assert (d["N"], d["M"]) == (208, 169)
    from parity_cases import ORACLE_TWIN
    for n in names:
        a, b = pmb.dims(n), orc.dims(ORACLE_TWIN.get(n, n))
        if n in ORACLE_TWIN:            # the drop-in classes travel as an opaque blob: NPARAM counts its 8-byte words
            a.pop("NPARAM"); b.pop("NPARAM")
        assert a == b
    with pytest.raises(PmbError):
        pmb.dims("no_such_problem")


def test_default_settings_match_reference(pmb):
    q = pmb.qp_default_settings()          # qp_base.hpp:17-53
    assert (q.max_iter, q.check_termination, q.adaptive_rho, q.adaptive_rho_interval) == (1000, 25, 0, 25)
    assert (q.rho, q.sigma, q.alpha, q.eps_abs, q.eps_rel) == (0.1, 1e-6, 1.0, 1e-3, 1e-3)
    s = pmb.sqp_default_settings()         # sqp_base.hpp:24-34
    assert (s.tau, s.eta, s.rho, s.eps_prim, s.eps_dual, s.max_iter, s.line_search_max_iter) == (0.5, 0.25, 0.5, 1e-3, 1e-3, 100, 100)
    q = pmb.sqp_default_qp_settings()      # sqp_base.hpp:83-90
    assert (q.max_iter, q.check_termination, q.adaptive_rho, q.adaptive_rho_interval, q.eps_abs, q.eps_rel) == (100, 10, 1, 50, 1e-4, 1e-4)


def test_struct_layouts():
    assert ctypes.sizeof(Dims) == 12 * 4
    assert ctypes.sizeof(SqpSettings) == 5 * 8 + 2 * 4
    assert ctypes.sizeof(QpSettings) == 2 * 8 + 4 * 4 + 3 * 8 + 2 * 4 + 8 + 2 * 4


def test_no_cpu_fallback(pmb):
    """Without a CUDA device every compute entry point must refuse (PMB_ERR_NO_DEVICE), never compute on the host."""
    if pmb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    import numpy as np
    with pytest.raises(PmbError, match="code -4|no CUDA device"):
        pmb.ocp("mobile_robot_6x2").cost(np.zeros(65), np.ones(1))
    with pytest.raises(PmbError):
        pmb.sqp("mobile_robot_6x2", 4)
    with pytest.raises(PmbError):
        pmb.bfgs_update(np.eye(2)[None], np.ones((1, 2)), np.ones((1, 2)))


def test_missing_library_fails_loudly(monkeypatch):
    import polympc_b200
    monkeypatch.setattr(polympc_b200, "_api", None)
    monkeypatch.setattr(polympc_b200, "LIB_PATH", "/nonexistent/libpolympc_b200.so")
    with pytest.raises(PmbError, match="no CPU fallback"):
        polympc_b200.load()


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under polympc_b200/ or include/ may reference it."""
    for base in ("polympc_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".hpp", ".h", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "liboracle" not in txt and "pyoracle" not in txt, os.path.join(dp, f)

"""The C++ host facade (include/polympc_b200.hpp) driven by a C++ transcription of the reference's MPCWrapperTest
(tests/control/mpc_wrapper_test.cpp:120-199)."""
import os
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_facade.cpp")


def _build(tmp_path, lib_dir, lib_name, extra):
    exe = str(tmp_path / "test_facade")
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include")] + extra + [SRC, "-o", exe, "-L" + lib_dir, "-l" + lib_name,
                                                                                       "-Wl,-rpath," + lib_dir]
    subprocess.run(cmd, check=True)
    return exe


def test_facade_on_warp_emulator(tmp_path, emu):
    """CPU: the facade over the warp-emulator build of the kernels (same sources, emu_ prefix)"""
    d = os.path.join(ROOT, "tests", "warp_emu")
    exe = _build(tmp_path, os.path.join(d, "_build"), "pmb_emu", ["-include", os.path.join(d, "emu_names.h")])
    out = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout


@pytest.mark.gpu
def test_facade_on_gpu(tmp_path, pmb):
    exe = _build(tmp_path, os.path.join(ROOT, "polympc_b200"), "polympc_b200", [])
    out = subprocess.run([exe, "64"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout and "sm_100a" in out.stdout

"""CPU suite: the product's kernel bodies, compiled by g++ against the lock-step warp emulator (tests/warp_emu, test
infrastructure), must reproduce the oracle bit for bit.  Small sizes — the emulator context-switches at every warp
synchronisation.  The same cases run against the real CUDA library in test_gpu_parity.py."""
import numpy as np
import pytest

import parity_cases as pc
from polympc_b200 import workloads as W


@pytest.mark.parametrize("name", ["mobile_robot_6x2", "mobile_robot_5x2", "mobile_robot_5x3", "cstr_5x2", "kite_4x2", "robot_obstacle_5x2",
                                  "dropin_robot_5x3", "dropin_cstr_5x2"])
def test_ocp_operators(emu, orc, name):
    pc.ocp_case(emu, orc, name, B=2, seed=1)


@pytest.mark.parametrize("N,M", [(1, 0), (2, 1), (5, 5), (12, 7), (33, 20), (40, 31)])
def test_qp(emu, orc, N, M):
    pc.qp_case(emu, orc, N, M, B=2, seed=N)


def test_qp_warm_start_and_rho_update(emu, orc):
    st = orc.qp_default_settings()
    st.max_iter = 200; st.adaptive_rho = 1; st.adaptive_rho_interval = 25; st.check_termination = 10; st.eps_abs = 1e-7; st.eps_rel = 1e-7
    r = pc.qp_case(emu, orc, 10, 6, B=3, seed=5, settings=st, warm=True, loose_rows=2)
    assert r["n_factor"].max() >= 2          # the rho re-factorisation path is exercised


def test_qp_max_iter_exceeded(emu, orc):
    st = orc.sqp_default_qp_settings(); st.max_iter = 7
    r = pc.qp_case(emu, orc, 6, 3, B=2, seed=9, settings=st)
    assert (r["info"]["iter"] == 8).all() and (r["info"]["status"] == 1).all()     # box_admm.hpp:199-202


@pytest.mark.parametrize("N", [2, 31, 65])
def test_bfgs(emu, orc, N):
    br = pc.bfgs_case(emu, orc, N)
    assert br[0] == 0 and br[2] == 2


def test_kkt_assemble(emu, orc):
    pc.kkt_case(emu, orc, 9, 5)


def test_sqp_robot(emu, orc):
    w = W.mobile_robot(2, seed=7, sqp_max_iter=10, ls_max_iter=10)
    ra, rb = pc.sqp_case(emu, orc, w)
    assert (rb["info"]["status"] == 0).any()


def test_sqp_cstr(emu, orc):
    w = W.cstr(1, seed=3, sqp_max_iter=4, ls_max_iter=10)
    pc.sqp_case(emu, orc, w)


def test_sqp_warm_restart_and_reset_guess(emu, orc):
    """second solve() warm-starts from the kept (x, lam); reset_guess() restores the guess given to set_primal/set_dual"""
    w = W.mobile_robot(2, seed=11, sqp_max_iter=2, ls_max_iter=10)
    outs = []
    for api in (emu, orc):
        s = api.sqp(w.name, 2); W.configure(s, w); s.solve()
        x1 = s.primal()
        s.set_initial_conditions(w.x0 + 0.01); s.solve()
        x2, l2 = s.primal(), s.dual()
        s.set_initial_conditions(w.x0); s.reset_guess(); s.solve()
        x3 = s.primal()
        outs.append((x1, x2, l2, x3))
        s.close()
    for a, b, n in zip(outs[0], outs[1], ("x1", "x2", "lam2", "x3")):
        pc.assert_same(a, b, n)
    pc.assert_same(outs[0][0], outs[0][3], "reset_guess reproduces the first solve")


def test_qp_nan_and_inf_data(emu, orc):
    """garbage in: NaN / inf entries must flow through both sides identically (Eigen maxCoeff pivot rule with NaN, IEEE inf)"""
    rng = np.random.default_rng(2)
    H, h, A, Alb, Aub, xlb, xub = pc.random_qp(rng, 4, 9, 5)
    H[0, 3, 3] = np.nan                   # NaN on the diagonal, not at the first pivot position
    H[1, 0, 0] = np.nan                   # NaN at the first position: stays there (maxCoeff starts from it)
    H[2, :, :] = np.nan                   # everything NaN
    H[3, 2, 2] = 1e308; h[3, 1] = 1e308   # overflow to inf
    st = orc.sqp_default_qp_settings(); st.max_iter = 30
    ra = emu.qp_solve(H, h, A, Alb, Aub, xlb, xub, st)
    rb = orc.qp_solve(H, h, A, Alb, Aub, xlb, xub, st)
    for k in ("perm", "ctype", "n_factor", "x", "y", "z", "q"):
        pc.assert_same(ra[k], rb[k], "qp." + k)
    for f in ("status", "iter"):
        pc.assert_same(ra["info"][f], rb["info"][f], "qp.info." + f)


def test_sqp_inequality_constraints(emu, orc):
    """NG = 1 (obstacle avoidance): inequality rows in the QP, lbg/ubg in the merit function and the termination test"""
    w = W.robot_obstacle(2, sqp_max_iter=6, ls_max_iter=10)
    pc.sqp_case(emu, orc, w)


@pytest.mark.parametrize("kind", ["robot", "cstr"])
def test_sqp_dropin_problem_classes(emu, orc, kind):
    """reference-style problem classes (Eigen functors over include/polympc_compat/, examples/dropin/*.hpp) driven through
    the kernels reproduce the hand-restated oracle model bit for bit — including the horizon the CSTR class sets in its
    own constructor"""
    import dataclasses
    w = W.mobile_robot(2, seed=5, grid="5x3", sqp_max_iter=6, ls_max_iter=10) if kind == "robot" else W.cstr(1, seed=4, sqp_max_iter=3, ls_max_iter=10)
    w = dataclasses.replace(w, name={"robot": "dropin_robot_5x3", "cstr": "dropin_cstr_5x2"}[kind])
    pc.sqp_case(emu, orc, w)
    if kind == "cstr":
        a, b = emu.ocp("dropin_cstr_5x2"), orc.ocp("cstr_5x2")
        b.set_time_limits(0.0, 100.0)
        pc.assert_same(a.time_nodes(), b.time_nodes(), "horizon from the class constructor")

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
    w = W.robot_obstacle(2, sqp_max_iter=3, ls_max_iter=10)          # (the GPU suite runs it to convergence)
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


def test_qp_nonfinite_jacobian_rows_with_zero_guess(emu, orc):
    """`m_z = A * x_guess` (box_admm.hpp:99) with the default zero guess is NaN — not 0 — for every row of A that holds a
    NaN or an infinity; that NaN decides later which multipliers of a diverged instance stay finite, and through the
    NaN-blind infinity norms whether SQPBase reports SOLVED (regression: the kernel used to start from z = 0)"""
    rng = np.random.default_rng(8)
    H, h, A, Alb, Aub, xlb, xub = pc.random_qp(rng, 3, 9, 5)
    A[0, 2, 4] = np.nan
    A[1, 0, 0] = np.inf; A[1, 4, 8] = -np.inf
    A[2, :, :] = np.nan; H[2, :, :] = np.nan; h[2, :] = np.nan; Alb[2, :] = np.nan; Aub[2, :] = np.nan; xlb[2, :] = np.nan; xub[2, :] = np.nan
    st = orc.sqp_default_qp_settings(); st.max_iter = 20
    ra = emu.qp_solve(H, h, A, Alb, Aub, xlb, xub, st)
    rb = orc.qp_solve(H, h, A, Alb, Aub, xlb, xub, st)
    for k in ("perm", "ctype", "n_factor", "x", "y", "z", "q"):
        pc.assert_same(ra[k], rb[k], "qp." + k)
    for f in ("status", "iter"):
        pc.assert_same(ra["info"][f], rb["info"][f], "qp.info." + f)


def test_sqp_cstr_warm_restart_that_diverges(emu, orc):
    """cstr_control_test.cpp:137-177 on the dense path: the warm-started second solve() linearises with the kept
    multipliers, its exact Hessian is indefinite and the iterates overflow to NaN after four iterations.  lpNorm<Infinity> of an
    all-NaN step is NaN, `NaN <= eps` is false, so the status stays MAX_ITER_EXCEEDED — on both sides, with identical traces"""
    outs = []
    for api in (emu, orc):
        # (the GPU suite runs the reference's 20 / 20; here 8 iterations are enough to pass the overflow at iteration 4)
        w = W.cstr(1, sqp_max_iter=8, ls_max_iter=20); w.x0[:] = [1.0, 0.5, 100.0, 100.0]
        s = api.sqp("cstr_5x2", 1); W.configure(s, w); s.set_trace(True); s.solve()
        first = s.info().copy()
        s.set_initial_conditions(np.array([[1.1, 0.508, 100.5, 100.1]])); s.solve()
        outs.append((first, s.info().copy(), s.primal(), s.dual(), s.stats(), s.trace(20)))
        s.close()
    a, b = outs
    assert a[0]["status"][0] == 0 and a[0]["iter"][0] == b[0]["iter"][0]
    assert not np.isfinite(b[2]).all() and b[1]["status"][0] == 1          # diverged and NOT reported as solved
    for f in ("iter", "qp_solver_iter", "status"):
        pc.assert_same(a[1][f], b[1][f], "warm.info." + f)
    for i, n in ((2, "x"), (3, "lam"), (4, "stats")):
        pc.assert_same(a[i], b[i], "warm." + n)
    for k in ("qp_iter", "ls_trials", "alpha", "bfgs", "qp_factor"):
        pc.assert_same(a[5][k], b[5][k], "warm.trace." + k)


@pytest.mark.parametrize("exact,gersh", [(0, 1), (1, 1)])            # (1, 0) runs in the GPU suite only: a minute on the emulator
def test_sqp_hessian_options(emu, orc, exact, gersh):
    """pmb_sqp_set_hessian_options — the SQPBase overrides of reference tests/control/minimal_time_test.cpp:90-135 as engine
    options: exact Hessian at every iteration, Gershgorin regularisation.  With the regulariser the warm-started CSTR solve
    of cstr_control_test.cpp converges for real (finite iterates) instead of overflowing."""
    outs = []
    for api in (emu, orc):
        w = W.cstr(1, sqp_max_iter=8 if (exact and not gersh) else 20, ls_max_iter=20); w.x0[:] = [1.0, 0.5, 100.0, 100.0]
        s = api.sqp("cstr_5x2", 1); W.configure(s, w); s.set_trace(True); s.set_hessian_options(exact, gersh); s.solve()
        first = s.info().copy()
        s.set_initial_conditions(np.array([[1.1, 0.508, 100.5, 100.1]])); s.solve()
        outs.append((first, s.info().copy(), s.primal(), s.dual(), s.stats(), s.trace(20)))
        s.close()
    a, b = outs
    for k in (0, 1):
        for f in ("iter", "qp_solver_iter", "status"):
            pc.assert_same(a[k][f], b[k][f], f"info[{k}]." + f)
    for i, n in ((2, "x"), (3, "lam"), (4, "stats")):
        pc.assert_same(a[i], b[i], n)
    for k in ("qp_iter", "ls_trials", "alpha", "bfgs", "qp_factor"):
        pc.assert_same(a[5][k], b[5][k], "trace." + k)
    if exact:
        assert (a[5]["bfgs"][a[5]["qp_iter"] > 0] == -1).all()          # no BFGS update was taken
    if gersh:
        assert a[1]["status"][0] == 0 and np.isfinite(a[2]).all() and a[1]["iter"][0] <= 5


def closed_loop(api, name, w, steps, dt):
    """MPC in closed loop (the caller of the path): solve, apply the first control to a unicycle plant
    (RK4 on the host, numpy), move the initial-condition bound to the new state, re-solve warm-started from the kept
    iterate — what `MPC::solve()` does when called once per control period (mpc_wrapper.hpp:89-99, 298)."""
    d = float(w.d[0])
    def f(x, u):
        return np.stack([u[:, 0] * np.cos(x[:, 2]) * np.cos(u[:, 1]), u[:, 0] * np.sin(x[:, 2]) * np.cos(u[:, 1]), u[:, 0] * np.sin(u[:, 1]) / d], axis=1)
    s = api.sqp(name, w.batch); W.configure(s, w)
    D = s.d
    x = w.x0.copy(); log = []
    for _ in range(steps):
        s.solve()
        var = s.primal()
        u0 = var[:, D["NX"] * D["NN"] + D["NU"] * (D["NN"] - 1):D["NX"] * D["NN"] + D["NU"] * D["NN"]]     # control at the initial time
        k1 = f(x, u0); k2 = f(x + 0.5 * dt * k1, u0); k3 = f(x + 0.5 * dt * k2, u0); k4 = f(x + dt * k3, u0)
        x = x + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        log.append((var.copy(), s.dual().copy(), s.info().copy(), x.copy()))
        s.set_initial_conditions(x)
    s.close()
    return log


def test_closed_loop_mpc(emu, orc):
    w = W.mobile_robot(2, seed=13, sqp_max_iter=4, ls_max_iter=10)
    la, lb = closed_loop(emu, w.name, w, 3, 0.1), closed_loop(orc, w.name, w, 3, 0.1)
    for k, (a, b) in enumerate(zip(la, lb)):
        pc.assert_same(a[0], b[0], f"step {k}: x"); pc.assert_same(a[1], b[1], f"step {k}: lam")
        for f in ("iter", "qp_solver_iter", "status"):
            pc.assert_same(a[2][f], b[2][f], f"step {k}: info." + f)
        pc.assert_same(a[3], b[3], f"step {k}: plant state")


def test_sqp_block_bfgs(emu, orc):
    """PMB_HESSIAN_BFGS_BLOCK: ContinuousOCP<..., SPARSE>::hessian_update_impl (continuous_ocp.hpp:2303-2431), the update every
    reference control test installs.  Plain and damped branches, robot and CSTR."""
    w = W.mobile_robot(2, seed=7, sqp_max_iter=10, ls_max_iter=10)
    ra, rb = pc.sqp_case(emu, orc, w, hessian_update=1)
    assert (rb["info"]["status"] == 0).any()
    w = W.cstr(1, seed=3, sqp_max_iter=6, ls_max_iter=10)
    ra, rb = pc.sqp_case(emu, orc, w, hessian_update=1)
    assert np.isfinite(rb["x"]).all()


@pytest.mark.parametrize("name", ["mobile_robot_5x2", "cstr_5x2"])
def test_block_bfgs_operator(emu, orc, name):
    br = pc.block_bfgs_case(emu, orc, name, B=3, seed=2)
    assert br[0] == 0 and br[1] == 1


def test_ocp_operators_with_optimised_parameter(emu, orc):
    """NP = 1 (reference ParkingOCP): parameter columns of the Jacobian, parameter gradient and the (., p) / (p, p) Hessian
    blocks accumulated over the nodes in the reference's loop order (continuous_ocp.hpp:860-872, 1314-1366, 2161-2172)"""
    pc.ocp_case(emu, orc, "parking_5x2", B=2, seed=4)


def test_sqp_minimal_time_parking(emu, orc):
    """the reference's tests/control/minimal_time_test.cpp:146-188 setup (free final time, exact Hessian at every iteration,
    Gershgorin regularisation, final-state box): SOLVED in fewer than max_iter iterations, like the reference asserts"""
    w = W.parking(1)
    w.sqp_max_iter = 4            # bit parity of the first iterations here; SOLVED in < 20 iterations is asserted on the oracle
    ra, rb = pc.sqp_case(emu, orc, w)   # (test_oracle_behaviour.py) and, against it, on the GPU
    assert np.isfinite(rb["x"]).all() and 0.0 < rb["x"][0, -1] < 10.0


@pytest.mark.parametrize("variant", [1, 2])
def test_ruiz_operators(emu, orc, variant):
    """RuizEquilibration::compute / unscale (qp_preconditioners.hpp:151-300, 364-404), DENSE and SPARSE variants, bit for bit"""
    pc.ruiz_case(emu, orc, 9, 5, B=4, seed=1, variant=variant)
    pc.ruiz_case(emu, orc, 33, 20, B=4, seed=2, variant=variant)
    pc.ruiz_case(emu, orc, 7, 0, B=2, seed=3, variant=variant)


def test_sqp_ruiz_and_filter_line_search(emu, orc):
    """SQPBase with a RuizEquilibration preconditioner (sqp_base.hpp:605-611, 662-667) and the filter line search of the
    reference's valet_parking_mpc_test.cpp:110-155; two solves, so the second one starts from the kept iterate AND the kept
    filter (max_depth 3 makes the oldest entry leave)"""
    w = W.mobile_robot(2, seed=11, sqp_max_iter=3, ls_max_iter=6)
    ra, rb = pc.sqp_case(emu, orc, w, preconditioner=2, line_search=1, filter_depth=3, solves=2)
    assert (rb["filter"][:, 0] >= 2).all()
    pc.sqp_case(emu, orc, w, preconditioner=1)


def test_sqp_factor_in_a_global_slot(emu, orc):
    """a problem whose packed factor does not fit in shared memory (kite 4 x 2: n = 261, 273 KB): the factor lives in a global
    slot and the 32 x 32 diagonal blocks are staged in shared memory for the substitutions (pmb_qp.hpp::ldlt_stage_diag_blocks;
    the last block is partial).  One SQP iteration = one QP with ~100 solves, bit for bit."""
    w = W.kite(1, grid="4x2", sqp_max_iter=1, ls_max_iter=4)
    pc.sqp_case(emu, orc, w)


@pytest.mark.parametrize("N,M", [(1, 0), (2, 1), (5, 5), (12, 7), (20, 13)])
def test_osqp_style_admm(emu, orc, N, M):
    """ADMM<N, M> of the reference (admm.hpp:112-213; KKT system of size 2N + M) as a batched operator, bit for bit"""
    pc.admm_case(emu, orc, N, M, B=3, seed=N)


def test_osqp_style_admm_adaptive_rho_relaxation_warm_start(emu, orc):
    st = orc.sqp_default_qp_settings(); st.adaptive_rho = 1; st.adaptive_rho_interval = 10; st.max_iter = 60; st.alpha = 1.6
    r = pc.admm_case(emu, orc, 9, 4, B=3, seed=5, settings=st, warm=True)
    assert (r["n_factor"] >= 2).any()


def test_sqp_with_osqp_style_admm(emu, orc):
    """SQPBase<..., ADMM<>> (pmb_sqp_set_qp_solver): the OSQP-style ADMM as the QP solver of the fused loop — alone and together
    with block BFGS, Ruiz equilibration and the filter line search, the combination valet_parking_mpc_test.cpp:168-172 aliases"""
    w = W.mobile_robot(2, seed=3, sqp_max_iter=3, ls_max_iter=6)
    pc.sqp_case(emu, orc, w, qp_solver=1)
    pc.sqp_case(emu, orc, w, qp_solver=1, preconditioner=2, line_search=1, hessian_update=1)

"""GPU suite: libpolympc_b200.so (sm_100a kernels) through its C ABI vs the CPU oracle, on a real B200.

Bar (BASELINE.json north_star): iterates within 1e-10 rel-inf of the CPU path and bit-exact active-set indices.  Because both
sides use the same deterministic elementary functions and summation orders we demand more: BIT-EXACT fp64 iterates,
multipliers and decision traces (tolerance 0).  Full-size runs (batch 8192) additionally check size-independent properties."""
import numpy as np
import pytest

import parity_cases as pc
from conftest import rel_inf
from polympc_b200 import workloads as W

pytestmark = pytest.mark.gpu

TOL_REL_INF = 1e-10      # the north_star tolerance; the assertions below are stricter (exact)


def test_device_present(pmb):
    assert pmb.device_count() >= 1
    assert "sm_100a" in pmb.version()


@pytest.mark.parametrize("name", ["mobile_robot_6x2", "mobile_robot_5x2", "mobile_robot_5x3", "cstr_5x2", "kite_4x2", "kite_12x1", "robot_obstacle_5x2",
                                  "parking_5x2", "dropin_robot_5x3", "dropin_cstr_5x2"])
def test_ocp_operators(pmb, orc, name):
    pc.ocp_case(pmb, orc, name, B=37, seed=1)


def test_ocp_against_casadi_golden(pmb, golden):
    """the CUDA transcription directly against the reference's CasADi fixtures (no oracle in between)"""
    o = pmb.ocp("mobile_robot_5x2")
    o.set_time_limits(0.0, 1.0)
    x, lam_eq = golden["x"], golden["lam"]
    B = x.shape[0]
    d = np.ones((B, 1))
    lam = np.concatenate([lam_eq, np.zeros((B, 55))], axis=1)
    r = o.lagrangian_gradient_hessian(x, lam, d)
    assert np.abs(r["cost"] - golden["cost"][:, 0]).max() <= 2e-14 * max(1.0, np.abs(golden["cost"]).max())
    assert np.abs(r["cost_grad"] - golden["cost_gradient"]).max() <= 2e-14
    assert np.abs(r["g"] - golden["constraints"]).max() <= 5e-14
    assert np.abs(r["jac"] - golden["constraints_jacobian"]).max() <= 5e-13
    assert np.abs(r["lag_grad"] - golden["lagrangian_gradient"]).max() <= 5e-13
    assert np.abs(r["hess"] - golden["lagrangian_hessian"]).max() <= 5e-13
    c, g, H = o.cost_gradient_hessian(x, d)
    assert np.abs(H - golden["cost_hessian"]).max() <= 2e-14


def test_empty_batch(pmb):
    o = pmb.ocp("mobile_robot_6x2")
    assert o.cost(np.zeros((0, 65)), np.zeros((0, 1))).shape == (0,)
    st = pmb.sqp_default_qp_settings()
    r = pmb.qp_solve(np.zeros((0, 3, 3)), np.zeros((0, 3)), np.zeros((0, 2, 3)), np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, 3)),
                     np.zeros((0, 3)), st)
    assert r["x"].shape == (0, 3)


@pytest.mark.parametrize("N,M", [(1, 0), (2, 1), (5, 5), (12, 7), (33, 20), (40, 31), (65, 39), (66, 44), (100, 92)])
def test_qp(pmb, orc, N, M):
    pc.qp_case(pmb, orc, N, M, B=33, seed=N)


def test_qp_reference_known_answers(pmb):
    st = pmb.qp_default_settings(); st.max_iter = 150        # box_admm_test.cpp:15-45
    r = pmb.qp_solve([[[4, 1], [1, 2]]], [[1, 1]], [[[1, 1]]], [[1]], [[1]], [[0, 0]], [[0.7, 0.7]], st)
    assert np.allclose(r["x"][0], [0.3, 0.7], rtol=1e-2) and r["info"]["status"][0] == 0 and r["info"]["iter"][0] < 150
    st = pmb.qp_default_settings(); st.max_iter = 200; st.adaptive_rho = 1; st.check_termination = 10   # :266-298
    r = pmb.qp_solve(np.zeros((1, 1, 1)), [[1.0]], np.zeros((1, 0, 1)), np.zeros((1, 0)), np.zeros((1, 0)), [[-1e6]], [[1e6]], st)
    assert np.allclose(r["x"][0], [-1e6], rtol=1e-2) and r["info"]["status"][0] == 0
    st.rho = 2                                                                                           # :300-334
    r = pmb.qp_solve(-np.ones((1, 1, 1)), [[0.0]], np.zeros((1, 0, 1)), np.zeros((1, 0)), np.zeros((1, 0)), [[-1.0]], [[2.0]], st,
                     x_guess=[[0.1]], y_guess=[[0.1]])
    assert np.allclose(r["x"][0], [2.0], rtol=1e-2) and r["info"]["status"][0] == 0


def test_qp_warm_start_and_rho_update(pmb, orc):
    st = orc.qp_default_settings()
    st.max_iter = 200; st.adaptive_rho = 1; st.adaptive_rho_interval = 25; st.check_termination = 10; st.eps_abs = 1e-7; st.eps_rel = 1e-7
    r = pc.qp_case(pmb, orc, 30, 16, B=65, seed=5, settings=st, warm=True, loose_rows=2)
    assert r["n_factor"].max() >= 2


def test_qp_max_iter_exceeded(pmb, orc):
    st = orc.sqp_default_qp_settings(); st.max_iter = 7
    r = pc.qp_case(pmb, orc, 6, 3, B=5, seed=9, settings=st)
    assert (r["info"]["iter"] == 8).all() and (r["info"]["status"] == 1).all()


@pytest.mark.parametrize("N", [2, 31, 65, 208])
def test_bfgs(pmb, orc, N):
    pc.bfgs_case(pmb, orc, N, B=9)


@pytest.mark.parametrize("N,M", [(9, 5), (65, 39)])
def test_kkt_assemble(pmb, orc, N, M):
    pc.kkt_case(pmb, orc, N, M, B=17)


def test_sqp_robot_vs_oracle(pmb, orc):
    """BASELINE config 2 inputs at a size the oracle finishes in seconds"""
    from oracle import pyoracle
    pyoracle.set_num_threads(8)
    w = W.mobile_robot(256)
    ra, rb = pc.sqp_case(pmb, orc, w)
    assert rel_inf(ra["x"], rb["x"]) <= TOL_REL_INF and rel_inf(ra["lam"], rb["lam"]) <= TOL_REL_INF
    assert (rb["info"]["status"] == 0).mean() > 0.5     # 10 / 20 iterations: most, not all, instances converge


def test_sqp_robot_reference_test_setup(pmb, orc):
    """mpc_wrapper_test.cpp:120-166 — x0 = (0.5, 0.5, 0.5), SQP 10/10: SOLVED (:145)"""
    w = W.mobile_robot(1, grid="5x3", sqp_max_iter=10, ls_max_iter=10)
    w.x0[:] = [0.5, 0.5, 0.5]
    ra, rb = pc.sqp_case(pmb, orc, w)
    assert ra["info"]["status"][0] == 0 and ra["info"]["iter"][0] <= 10


def test_sqp_cstr_vs_oracle(pmb, orc):
    w = W.cstr(128)
    ra, rb = pc.sqp_case(pmb, orc, w)
    assert (ra["info"]["status"] == 0).all()          # cstr_control_test.cpp:177 asserts SOLVED


def test_sqp_kite_vs_oracle(pmb, orc):
    w = W.kite(16, grid="4x2", sqp_max_iter=6, ls_max_iter=10)
    pc.sqp_case(pmb, orc, w)
    w = W.kite(8, sqp_max_iter=3, ls_max_iter=10)
    pc.sqp_case(pmb, orc, w)


def test_sqp_full_size_properties(pmb, orc):
    """BASELINE config 2 at full size (batch 8192): properties that need no oracle run of the whole batch"""
    B = 8192
    w = W.mobile_robot(B)
    r = pc.solve_workload(pmb, w)
    info = r["info"]
    assert (info["status"] == 0).mean() > 0.95
    assert info["iter"].min() >= 1 and info["iter"].max() <= 100
    x = r["x"]
    assert np.isfinite(x).all() and np.isfinite(r["lam"]).all()
    # initial condition = box equality on the last state block (mpc_wrapper.hpp:89-93): satisfied to QP tolerance
    solved = info["status"] == 0
    assert np.abs(x[solved, 36:39] - w.x0[solved]).max() < 1e-3
    # control bounds hold up to the ADMM tolerance
    u = x[:, 39:].reshape(B, 13, 2)
    assert (np.abs(u[solved, :, 0]).max() <= 1.5 + 1e-3) and (np.abs(u[solved, :, 1]).max() <= 0.75 + 1e-3)
    # collocation residuals of converged instances: re-evaluated by the independent equalities operator
    o = pmb.ocp(w.name); o.set_time_limits(w.t0, w.tf)
    c = o.equalities(x, np.full((B, 1), 2.0))
    assert np.abs(c[solved]).max() <= 1e-3 + 1e-12       # termination criterion max_viol <= eps_prim (sqp_base.hpp:524-528)
    # batch-order independence + determinism: a shuffled sub-batch reproduces the same iterates bit for bit
    idx = np.random.default_rng(0).permutation(B)[:512]
    w2 = W.mobile_robot(B); w2.x0 = w.x0[idx]
    r2 = pc.solve_workload(pmb, w2)
    pc.assert_same(r2["x"], x[idx], "shuffled sub-batch x")
    pc.assert_same(r2["info"]["iter"], info["iter"][idx], "shuffled sub-batch iter")
    # and a random sample of the full batch against the oracle
    sub = idx[:64]
    w3 = W.mobile_robot(B); w3.x0 = w.x0[sub]
    rb = pc.solve_workload(orc, w3)
    pc.assert_same(x[sub], rb["x"], "sample vs oracle x")
    pc.assert_same(r["lam"][sub], rb["lam"], "sample vs oracle lam")


def test_sqp_inequality_constraints_vs_oracle(pmb, orc):
    """NG = 1 (obstacle avoidance): bit-exact vs oracle, and the constraint holds on converged instances"""
    w = W.robot_obstacle(192)
    ra, rb = pc.sqp_case(pmb, orc, w)
    solved = ra["info"]["status"] == 0
    assert solved.mean() > 0.7
    X = ra["x"][:, :33].reshape(-1, 11, 3)
    g = (X[:, :, 0] - 0.25) ** 2 + (X[:, :, 1] - 0.25) ** 2
    assert g[solved].min() >= 0.09 - 1e-3


def test_cstr_full_size_properties(pmb, orc):
    """BASELINE config 3 at full size (CSTR 5x2, batch 4096)"""
    B = 4096
    w = W.cstr(B)
    r = pc.solve_workload(pmb, w)
    info = r["info"]
    assert (info["status"] == 0).mean() > 0.99            # cstr_control_test.cpp:177 asserts SOLVED
    solved = info["status"] == 0
    x = r["x"]
    assert np.isfinite(x).all()
    assert np.abs(x[solved, 40:44] - w.x0[solved]).max() < 1e-2 * 100         # initial condition (states ~100) to QP tolerance
    u = x[:, 44:].reshape(B, 11, 2)
    assert u[solved, :, 0].min() >= 3.0 - 1e-2 and u[solved, :, 0].max() <= 35.0 + 1e-2
    assert u[solved, :, 1].min() >= -9000.0 - 1.0 and u[solved, :, 1].max() <= 0.0 + 1.0
    sub = np.random.default_rng(1).permutation(B)[:48]
    w2 = W.cstr(B); w2.x0 = w.x0[sub]
    rb = pc.solve_workload(orc, w2)
    pc.assert_same(x[sub], rb["x"], "sample vs oracle x")
    pc.assert_same(info["iter"][sub], rb["info"]["iter"], "sample vs oracle iter")


def test_kite_full_size_properties(pmb, orc):
    """BASELINE config 4 at full size (kite 12x1, batch 1024; factor in the L2-resident global slot)"""
    B = 1024
    w = W.kite(B)
    r = pc.solve_workload(pmb, w)
    info = r["info"]
    assert (info["status"] == 0).mean() > 0.95
    x = r["x"]
    assert np.isfinite(x).all()
    solved = info["status"] == 0
    assert np.abs(x[solved, 156:169] - w.x0[solved]).max() < 1e-2             # initial condition on the last state block
    sub = np.random.default_rng(2).permutation(B)[:8]
    w2 = W.kite(B); w2.x0 = w.x0[sub]
    rb = pc.solve_workload(orc, w2)
    pc.assert_same(x[sub], rb["x"], "sample vs oracle x")
    pc.assert_same(r["lam"][sub], rb["lam"], "sample vs oracle lam")


def test_warm_restart_matches_oracle(pmb, orc):
    """MPC re-solve: a second solve() warm-starts from the kept (x, lam) (mpc_wrapper.hpp / sqp_base.hpp:568-696).
    One instance of this batch re-linearises to an indefinite exact Hessian and boxADMM diverges to inf/NaN on BOTH sides
    (the reference has no safeguard either); iterates and decision traces still agree there, the multipliers of such a
    non-finite instance are only required to be non-finite garbage on both sides ('non-finite instances')."""
    w = W.mobile_robot(32, sqp_max_iter=3, ls_max_iter=10)
    outs = []
    for api in (pmb, orc):
        s = api.sqp(w.name, 32); W.configure(s, w); s.set_trace(True); s.solve()
        s.set_initial_conditions(w.x0 + 0.01); s.solve()
        outs.append((s.primal(), s.dual(), s.info(), s.trace(3)))
        s.close()
    pc.assert_same(outs[0][0], outs[1][0], "x")
    pc.assert_same(outs[0][2]["iter"], outs[1][2]["iter"], "iter")
    for k in ("qp_iter", "alpha", "bfgs", "ls_trials", "qp_factor"):
        pc.assert_same(outs[0][3][k], outs[1][3][k], "trace." + k)
    finite = np.isfinite(outs[1][0]).all(axis=1)
    assert finite.sum() >= 30
    pc.assert_same(outs[0][1][finite], outs[1][1][finite], "lam (finite instances)")
    if (~finite).any():
        assert not np.isfinite(outs[0][1][~finite]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,batch", [("robot", 512), ("cstr", 256)])
def test_sqp_dropin_problem_classes(pmb, orc, kind, batch):
    """reference-style problem classes (examples/dropin/*.hpp: Eigen functors over include/polympc_compat/) on the GPU ==
    the fixture-pinned oracle model, bit for bit, and == the engine's hand-written twin problem"""
    import dataclasses
    w = W.mobile_robot(batch, seed=21, grid="5x3", sqp_max_iter=10, ls_max_iter=10) if kind == "robot" else W.cstr(batch, seed=22, sqp_max_iter=20, ls_max_iter=20)
    twin = w.name
    wd = dataclasses.replace(w, name={"robot": "dropin_robot_5x3", "cstr": "dropin_cstr_5x2"}[kind])
    ra, rb = pc.sqp_case(pmb, orc, wd)
    rt = pc.solve_workload(pmb, w)
    pc.assert_same(ra["x"], rt["x"], "drop-in class vs hand-written twin on the GPU")
    assert (rb["info"]["status"] == 0).mean() > 0.5     # 10 / 20 iterations: most, not all, instances converge


def test_qp_nonfinite_jacobian_rows_with_zero_guess(pmb, orc):
    """see tests/test_emu_parity.py::test_qp_nonfinite_jacobian_rows_with_zero_guess"""
    import test_emu_parity as te
    te.test_qp_nonfinite_jacobian_rows_with_zero_guess(pmb, orc)


def test_sqp_cstr_warm_restart_that_diverges(pmb, orc):
    import test_emu_parity as te
    te.test_sqp_cstr_warm_restart_that_diverges(pmb, orc)


@pytest.mark.parametrize("exact,gersh", [(1, 0), (0, 1), (1, 1)])
def test_sqp_hessian_options(pmb, orc, exact, gersh):
    import test_emu_parity as te
    te.test_sqp_hessian_options(pmb, orc, exact, gersh)
    w = W.mobile_robot(256, seed=31, sqp_max_iter=15, ls_max_iter=15)
    outs = []
    for api in (pmb, orc):
        s = api.sqp(w.name, w.batch); W.configure(s, w); s.set_hessian_options(exact, gersh); s.solve()
        outs.append((s.primal(), s.dual(), s.info().copy())); s.close()
    pc.assert_same(outs[0][0], outs[1][0], "x"); pc.assert_same(outs[0][1], outs[1][1], "lam")
    for f in ("iter", "qp_solver_iter", "status"):
        pc.assert_same(outs[0][2][f], outs[1][2][f], f)


def test_closed_loop_mpc(pmb, orc):
    """20 control periods of 256 robots in closed loop, warm-started re-solves: every iterate of every period bit-identical
    to the oracle, the fleet approaches the origin"""
    import test_emu_parity as te
    w = W.mobile_robot(256, seed=17, sqp_max_iter=10, ls_max_iter=10)
    la, lb = te.closed_loop(pmb, w.name, w, 20, 0.1), te.closed_loop(orc, w.name, w, 20, 0.1)
    for k, (a, b) in enumerate(zip(la, lb)):
        pc.assert_same(a[0], b[0], f"step {k}: x"); pc.assert_same(a[1], b[1], f"step {k}: lam")
        for f in ("iter", "qp_solver_iter", "status"):
            pc.assert_same(a[2][f], b[2][f], f"step {k}: info." + f)
    fin = np.isfinite(la[-1][3]).all(axis=1)
    assert fin.mean() > 0.9        # a few warm restarts linearise to an indefinite Hessian and diverge (on both sides, identically)
    assert np.median(np.linalg.norm(la[-1][3][fin, :2], axis=1)) < np.median(np.linalg.norm(w.x0[fin, :2], axis=1))


@pytest.mark.parametrize("name", ["mobile_robot_6x2", "cstr_5x2", "parking_5x2", "kite_4x2"])
def test_block_bfgs_operator(pmb, orc, name):
    """ContinuousOCP<..., SPARSE>::hessian_update_impl (continuous_ocp.hpp:2303-2431) as an operator, plain and damped branch"""
    br = pc.block_bfgs_case(pmb, orc, name, B=33, seed=2)
    assert br[0] == 0 and br[1] == 1


@pytest.mark.parametrize("kind,batch", [("mobile_robot", 1024), ("cstr", 512)])
def test_sqp_block_bfgs_vs_oracle(pmb, orc, kind, batch):
    """PMB_HESSIAN_BFGS_BLOCK — the SPARSE semantics every reference control test asks for — bit-exact against the oracle"""
    w = W.WORKLOADS[kind](batch, sqp_max_iter=20, ls_max_iter=20)
    ra, rb = pc.sqp_case(pmb, orc, w, hessian_update=1)
    assert np.isfinite(rb["x"]).all() and (rb["info"]["status"] == 0).mean() > 0.9


def test_sqp_minimal_time_parking_vs_oracle(pmb, orc):
    """NP = 1, exact Hessian at every iteration, Gershgorin regularisation, final-state box: the setup of the reference's
    tests/control/minimal_time_test.cpp:146-188 on 256 initial states; instance 0 is the reference's own (SOLVED, iter < 20)"""
    w = W.parking(256)
    ra, rb = pc.sqp_case(pmb, orc, w)
    assert rb["info"]["status"][0] == 0 and rb["info"]["iter"][0] < w.sqp_max_iter
    assert (rb["info"]["status"] == 0).mean() > 0.85


def test_sqp_codegen_test_setup(pmb, orc):
    """tests/solvers/sqp/codegen_test.cpp:402-437: robot 5 x 2 on [0, 1], d = 1, exact Hessian at every iteration, SQP 10 / 10, QP
    max_iter 1000, x0 = (0.5, 0.5, 0.5): the reference asserts SOLVED and iter < 10"""
    w = W.mobile_robot(1, grid="5x2", sqp_max_iter=10, ls_max_iter=10)
    w.t0, w.tf, w.d = 0.0, 1.0, np.array([1.0]); w.x0[:] = [0.5, 0.5, 0.5]; w.exact_hessian = True
    outs = []
    for api in (pmb, orc):
        s = api.sqp(w.name, 1); W.configure(s, w)
        q = s.qp_settings(); q.max_iter = 1000; s.set_qp_settings(q)
        s.solve(); outs.append((s.primal(), s.dual(), s.info())); s.close()
    pc.assert_same(outs[0][0], outs[1][0], "x"); pc.assert_same(outs[0][1], outs[1][1], "lam")
    assert outs[1][2]["status"][0] == 0 and outs[1][2]["iter"][0] < 10


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("N,M", [(9, 5), (7, 0), (65, 39), (80, 48), (208, 169)])
def test_ruiz_operators(pmb, orc, N, M, variant):
    """RuizEquilibration::compute / unscale (qp_preconditioners.hpp:151-300, 364-404) as batched operators, bit for bit"""
    pc.ruiz_case(pmb, orc, N, M, B=16, seed=N + variant, variant=variant)


@pytest.mark.parametrize("pre,ls", [(1, 0), (2, 0), (0, 1), (2, 1)])
def test_sqp_preconditioner_and_line_search_vs_oracle(pmb, orc, pre, ls):
    """RuizEquilibration (DENSE / SPARSE) around every QP and the filter line search, robot 6 x 2, two solves (kept filter)"""
    w = W.mobile_robot(512, sqp_max_iter=20, ls_max_iter=20)
    ra, rb = pc.sqp_case(pmb, orc, w, preconditioner=pre, line_search=ls, filter_depth=4, solves=2)
    assert np.isfinite(rb["x"]).all() and (rb["info"]["status"] == 0).mean() > 0.9


def test_sqp_ruiz_cstr_vs_oracle(pmb, orc):
    """the CSTR's badly scaled QPs (states ~100, inputs ~1e4) are what equilibration is for"""
    w = W.cstr(256, sqp_max_iter=20, ls_max_iter=20)
    ra, rb = pc.sqp_case(pmb, orc, w, hessian_update=1, preconditioner=2)
    assert np.isfinite(rb["x"]).all()


def test_valet_parking_test_setup_vs_oracle(pmb, orc):
    """tests/control/valet_parking_mpc_test.cpp:175-235 on 256 pairs of pinned states; pair 0 is the reference's own
    ((0.5, 0.5, 0.5) then (0.3, 0.4, 0.45): SOLVED, iter < 10 both times)"""
    rng = np.random.default_rng(5)
    a = np.column_stack([rng.uniform(-1, 1, 256), rng.uniform(-1, 1, 256), rng.uniform(-0.7, 0.7, 256)]); a[0] = [0.5, 0.5, 0.5]
    b = a + rng.uniform(-0.2, 0.2, a.shape); b[0] = [0.3, 0.4, 0.45]
    ra, rb = pc.valet_parking_case(pmb, orc, a, b)
    for r in rb:
        assert r["info"]["status"][0] == 0 and r["info"]["iter"][0] < 10
    assert (rb[0]["info"]["status"] == 0).mean() > 0.8


@pytest.mark.parametrize("kind,batch", [("mobile_robot", 8192), ("cstr", 4096)])
def test_sqp_full_batch_bit_exact_vs_oracle(pmb, orc, kind, batch):
    """BASELINE configs 2 and 3 at FULL size, every instance against the oracle: iterates, multipliers, status, iteration counts
    and the whole decision trace (ADMM trips, factorisations, BFGS branch, line-search trials, alpha per SQP iteration) are
    identical for all 8192 / 4096 instances (exact arithmetic).  The oracle needs a few seconds on the box's host cores."""
    import os
    from oracle import pyoracle
    pyoracle.set_num_threads(os.cpu_count() or 8)
    w = W.WORKLOADS[kind](batch)
    ra, rb = pc.sqp_case(pmb, orc, w)
    assert ra["x"].shape[0] == batch
    assert (rb["info"]["status"] == 0).mean() > 0.95


@pytest.mark.parametrize("N,M", [(1, 0), (2, 1), (5, 5), (12, 7), (33, 20), (65, 39), (80, 48), (100, 56)])
def test_osqp_style_admm(pmb, orc, N, M):
    """ADMM<N, M> of the reference (admm.hpp:112-213) as a batched operator: KKT systems of size 2N + M up to 256 (factor in shared
    memory up to 192, in a global slot beyond), bit for bit — incl. the robot's (65, 39) and the 5 x 3 grid's (80, 48)"""
    pc.admm_case(pmb, orc, N, M, B=24, seed=N + 1)


def test_osqp_style_admm_adaptive_rho_relaxation_warm_start(pmb, orc):
    st = orc.sqp_default_qp_settings(); st.adaptive_rho = 1; st.adaptive_rho_interval = 10; st.max_iter = 200; st.alpha = 1.6
    r = pc.admm_case(pmb, orc, 20, 9, B=64, seed=5, settings=st, warm=True)
    assert (r["n_factor"] >= 2).any() and (r["info"]["status"] == 0).any()


def test_osqp_style_admm_reference_known_answer(pmb):
    """admm_solver_test.cpp:16-45 (admmSimpleQP) on the GPU: (0.3, 0.7) to 1e-2, SOLVED, iter < 1000"""
    st = pmb.qp_default_settings(); st.max_iter = 1000
    r = pmb.qp_solve_admm(np.array([[[4.0, 1.0], [1.0, 2.0]]]), [[1.0, 1.0]], np.array([[[1.0, 1.0]]]), [[1.0]], [[1.0]], [[0.0, 0.0]], [[0.7, 0.7]], st)
    assert np.allclose(r["x"][0], [0.3, 0.7], rtol=1e-2) and r["info"]["status"][0] == 0 and r["info"]["iter"][0] < 1000


@pytest.mark.parametrize("kind,batch", [("mobile_robot", 1024), ("cstr", 256), ("robot_obstacle", 256)])
def test_sqp_with_osqp_style_admm_vs_oracle(pmb, orc, kind, batch):
    """SQPBase<..., ADMM<>>: the OSQP-style ADMM (KKT systems of size 2N + M = 169 / 176 / 176) as the QP solver of the fused loop"""
    w = W.WORKLOADS[kind](batch, sqp_max_iter=20, ls_max_iter=20)
    ra, rb = pc.sqp_case(pmb, orc, w, qp_solver=1)
    assert np.isfinite(rb["x"]).all() and (rb["info"]["status"] == 0).mean() > (0.5 if kind == "robot_obstacle" else 0.8)   # 20 / 20 iterations


def test_sqp_with_osqp_style_admm_and_the_valet_options(pmb, orc):
    """ADMM<> + block BFGS + RuizEquilibration<SPARSE> + filter line search on the valet-parking setup (5 x 3 grid, 2N + M = 208)"""
    rng = np.random.default_rng(6)
    a = np.column_stack([rng.uniform(-1, 1, 128), rng.uniform(-1, 1, 128), rng.uniform(-0.7, 0.7, 128)]); a[0] = [0.5, 0.5, 0.5]
    b = a + rng.uniform(-0.2, 0.2, a.shape); b[0] = [0.3, 0.4, 0.45]
    ra, rb = pc.valet_parking_case(pmb, orc, a, b, qp_solver=1)
    assert rb[0]["info"]["status"][0] == 0 and rb[1]["info"]["status"][0] == 0


def test_osqp_style_admm_limits(pmb):
    """exact arithmetic only; instantiated up to 2N + M = 256 (the kite's 585 is refused with PMB_ERR_UNSUPPORTED)"""
    from polympc_b200.capi import PmbError
    s = pmb.sqp("kite_12x1", 2)
    with pytest.raises(PmbError):
        s.set_qp_solver(1)
    s.close()
    w = W.mobile_robot(4)
    s = pmb.sqp(w.name, 4); W.configure(s, w); s.set_qp_solver(1); s.set_arithmetic(1)
    with pytest.raises(PmbError):
        s.solve()
    s.set_arithmetic(0); s.solve()
    assert (s.info()["status"] == 0).all()
    s.close()

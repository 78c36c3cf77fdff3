"""Behavioural pins of the reference's control tests, checked on the oracle."""
import numpy as np

import parity_cases as pc
from polympc_b200 import workloads as W


def test_robot_mpc_wrapper_setup_solves(orc):
    """mpc_wrapper_test.cpp:120-166: 5x3 grid, t in [0,2], d=2, x0=(.5,.5,.5), SQP 10/10 -> SOLVED, iter <= max_iter"""
    w = W.mobile_robot(1, grid="5x3", sqp_max_iter=10, ls_max_iter=10)
    w.x0[:] = [0.5, 0.5, 0.5]
    r = pc.solve_workload(orc, w)
    assert r["info"]["status"][0] == 0 and 1 <= r["info"]["iter"][0] <= 10
    x = r["x"][0]
    assert np.abs(x[45:48] - 0.5).max() < 1e-3                 # initial condition on the LAST state block
    u = x[48:].reshape(16, 2)
    assert np.abs(u[:, 0]).max() <= 1.5 + 1e-3 and np.abs(u[:, 1]).max() <= 0.75 + 1e-3


def test_cstr_setup_solves(orc):
    """cstr_control_test.cpp:137-177: SQP 20/20 -> SOLVED"""
    w = W.cstr(1, sqp_max_iter=20, ls_max_iter=20)
    w.x0[:] = [1.0, 0.5, 100.0, 100.0]
    r = pc.solve_workload(orc, w)
    assert r["info"]["status"][0] == 0 and r["info"]["iter"][0] <= 20


def test_admm_trip_counts_are_multiples_of_check_interval(orc):
    w = W.mobile_robot(8, sqp_max_iter=10, ls_max_iter=10)
    r = pc.solve_workload(orc, w)
    q = r["trace"]["qp_iter"]
    done = q[q > 0]
    assert ((done % 10 == 0) | (done == 101)).all()


def test_info_iter_counts_qps(orc):
    w = W.mobile_robot(4, sqp_max_iter=3, ls_max_iter=10)
    r = pc.solve_workload(orc, w)
    assert (r["info"]["iter"] <= 3).all() and (r["info"]["iter"] >= 1).all()
    n_qp = (r["trace"]["qp_iter"] > 0).sum(axis=1)
    assert np.array_equal(n_qp, r["info"]["iter"])
    assert np.array_equal(r["info"]["qp_solver_iter"], np.where(r["trace"]["qp_iter"] > 0, r["trace"]["qp_iter"], 0).sum(axis=1))


def test_codegen_test_setup(orc):
    """tests/solvers/sqp/codegen_test.cpp:402-437 (the reference's own SQP pin on the CasADi-fixture problem): robot 5 x 2 on
    [0, 1], d = 1, exact Hessian at every iteration, SQP 10 / 10, QP max_iter 1000, x0 = (0.5, 0.5, 0.5) -> SOLVED, iter < 10"""
    import numpy as np
    from polympc_b200 import workloads as W
    w = W.mobile_robot(1, grid="5x2", sqp_max_iter=10, ls_max_iter=10)
    w.t0, w.tf, w.d = 0.0, 1.0, np.array([1.0]); w.x0[:] = [0.5, 0.5, 0.5]; w.exact_hessian = True
    s = orc.sqp(w.name, 1); W.configure(s, w)
    q = s.qp_settings(); q.max_iter = 1000; s.set_qp_settings(q)
    s.solve()
    info = s.info(); s.close()
    assert info["status"][0] == 0 and info["iter"][0] < 10


def test_minimal_time_setup(orc):
    """tests/control/minimal_time_test.cpp:146-188: SOLVED, iter < max_iter (20), final time inside its box"""
    from polympc_b200 import workloads as W
    w = W.parking(1)
    s = orc.sqp(w.name, 1); W.configure(s, w); s.solve()
    info, x = s.info(), s.primal(); s.close()
    assert info["status"][0] == 0 and info["iter"][0] < 20 and 0.0 < x[0, -1] < 10.0


def test_valet_parking_test_setup(orc):
    """tests/control/valet_parking_mpc_test.cpp:175-235 (filter line search with LSFilter, block BFGS, RuizEquilibration<SPARSE>):
    the reference asserts SOLVED and iter < max_iter for the cold solve and for the warm-started one"""
    import parity_cases as pc
    first, second = pc.valet_parking_solve(orc, [0.5, 0.5, 0.5], [0.3, 0.4, 0.45])
    for r in (first, second):
        assert r["info"]["status"][0] == 0 and r["info"]["iter"][0] < 10
    assert second["filter"][0, 0] >= 1                       # the filter lives on between the two solves
    # the dense variant of the preconditioner and the identity solve the same problem
    for pre in (0, 1):
        a, b = pc.valet_parking_solve(orc, [0.5, 0.5, 0.5], [0.3, 0.4, 0.45], preconditioner=pre)
        assert a["info"]["status"][0] == 0 and b["info"]["status"][0] == 0
        assert np.abs(a["x"] - first["x"]).max() < 1e-2


def test_ruiz_equilibration_properties(orc):
    """RuizEquilibration::compute (qp_preconditioners.hpp:151-300): after <= 4 passes the row / column infinity norms of the
    scaled [H A'; A 0] are within a factor of a few of each other, zero rows / columns are left alone, the scaling is undone by
    unscale() to rounding, and infinite box bounds stay infinite"""
    import parity_cases as pc
    for variant in (1, 2):
        r = pc.ruiz_case(orc, orc, 20, 11, B=4, seed=5, variant=variant)
        assert np.isfinite(r["D"]).all() and np.isfinite(r["E"]).all() and (r["c"] > 0).all()
        assert np.isinf(r["l"][:, 0]).all() and np.isinf(r["u"][:, 1]).all()
        assert r["D"][1, 10] == 1.0                           # the empty column keeps scale 1 (the zero guard)
        coln = np.maximum(np.abs(r["H"][0]).max(axis=0) / r["c"][0], np.abs(r["A"][0]).max(axis=0))
        assert coln.max() / coln.min() < 30.0


def test_osqp_style_admm_known_answers(orc):
    """tests/solvers/qp/admm_solver_test.cpp: admmSimpleQP (:16-45), admmConstraintViolation (:114-151), admmSimpleLP (:303-334),
    admmNonConvex (:336-371), admmAdaptiveRho-style settings — the reference's own known answers for ADMM<>"""
    st = lambda **kw: _qp_settings(orc, **kw)
    H = np.array([[[4.0, 1.0], [1.0, 2.0]]]); h = np.array([[1.0, 1.0]]); A = np.array([[[1.0, 1.0]]])
    r = orc.qp_solve_admm(H, h, A, [[1.0]], [[1.0]], [[0.0, 0.0]], [[0.7, 0.7]], st(max_iter=1000))
    assert np.allclose(r["x"][0], [0.3, 0.7], rtol=1e-2) and r["info"]["status"][0] == 0 and r["info"]["iter"][0] < 1000
    r = orc.qp_solve_admm(H, h, A, [[1.0]], [[1.0]], [[0.0, 0.0]], [[0.7, 0.7]], st(eps_rel=1e-4, eps_abs=1e-4))
    x = r["x"][0]
    assert min(x.sum() - 1.0, x.min()) >= -1e-3 and max(x.sum() - 1.0, (x - 0.7).max()) <= 1e-3
    none = (np.zeros((1, 0, 1)), np.zeros((1, 0)), np.zeros((1, 0)))
    r = orc.qp_solve_admm(np.array([[[0.0]]]), [[1.0]], none[0], none[1], none[2], [[-1e6]], [[1e6]],
                          st(max_iter=200, alpha=1.0, adaptive_rho=1, check_termination=10))
    assert abs(r["x"][0, 0] + 1e6) <= 1e-2 * 1e6 and r["info"]["status"][0] == 0 and r["info"]["iter"][0] < 200
    r = orc.qp_solve_admm(np.array([[[-1.0]]]), [[0.0]], none[0], none[1], none[2], [[-1.0]], [[2.0]],
                          st(max_iter=200, alpha=1.0, adaptive_rho=1, rho=2.0, check_termination=10), x_guess=[[0.1]], y_guess=[[0.1]])
    assert abs(r["x"][0, 0] - 2.0) <= 2e-2 and r["info"]["status"][0] == 0 and r["info"]["iter"][0] < 200
    # same optimum as boxADMM on a random strictly convex QP
    import parity_cases as pc
    rng = np.random.default_rng(3)
    q = pc.random_qp(rng, 4, 8, 3)
    s = st(max_iter=4000, eps_abs=1e-8, eps_rel=1e-8, check_termination=10)
    a, b = orc.qp_solve_admm(*q, s), orc.qp_solve(*q, s)
    assert (a["info"]["status"] == 0).all() and np.abs(a["x"] - b["x"]).max() < 1e-5


def _qp_settings(orc, **kw):
    s = orc.qp_default_settings()
    for k, v in kw.items():
        setattr(s, k, v)
    return s

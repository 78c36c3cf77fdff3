"""Pins the CPU oracle (oracle/, test infrastructure) against what the reference itself provides:
  * its CasADi-generated C fixtures of the 5x2 mobile-robot NLP (tests/solvers/sqp/casadi_codegen/*.cpp) through the
    committed golden vectors tests/golden/casadi_robot_5x2.npz (generator: tests/golden/make_casadi_golden.py);
  * the known answers of its QP / BFGS / classification unit tests;
  * the behavioural pins (SOLVED, iteration bounds) of its control tests.
"""
import numpy as np
import pytest

from polympc_b200 import capi


# ---- Chebyshev tables: constants visible in the CasADi fixture / SURVEY App. D --------------------------------------
def test_cheb_tables_known_values(orc):
    nodes, D, w = orc.cheb_tables(5)
    assert abs(w[0] - 0.04) < 1e-16 and abs(w[1] - 0.36074304120001122) < 1e-15 and abs(w[2] - 0.59925695879998886) < 1e-15
    assert abs(D[0, 0] - 8.5) < 1e-13 and abs(D[0, 1] + 10.472135954999581) < 1e-13
    nodes, D, w = orc.cheb_tables(6)
    assert abs(w[0] - 1.0 / 35.0) < 1e-16 and abs(D[0, 0] - 73.0 / 6.0) < 1e-13 and abs(D[0, 1] + 14.928203230275516) < 1e-13
    nodes, D, w = orc.cheb_tables(12)
    assert abs(w[0] - 1.0 / 143.0) < 1e-16 and abs(D[0, 0] - 289.0 / 6.0) < 1e-12


@pytest.mark.parametrize("P", [2, 3, 4, 5, 6, 7, 12])
def test_cheb_tables_properties(orc, P):
    nodes, D, w = orc.cheb_tables(P)
    assert np.allclose(nodes, np.cos(np.arange(P + 1) * np.pi / P), atol=1e-15)
    assert abs(w.sum() - 2.0) < 1e-14                       # integrates 1 exactly
    assert abs(w @ nodes ** 2 - 2.0 / 3.0) < 1e-14          # and x^2
    assert np.abs(D @ np.ones(P + 1)).max() < 1e-12         # derivative of a constant
    assert np.abs(D @ nodes ** 2 - 2 * nodes).max() < 1e-11   # derivative of x^2 is exact for P >= 2


# ---- transcription vs the reference's CasADi fixtures -------------------------------------------------------------------
@pytest.fixture(scope="module")
def robot5x2(orc):
    o = orc.ocp("mobile_robot_5x2")
    o.set_time_limits(0.0, 1.0)
    return o


def test_casadi_cost_gradient_hessian(robot5x2, golden):
    x = golden["x"]
    d = np.ones((x.shape[0], 1))
    c, g, H = robot5x2.cost_gradient_hessian(x, d)
    assert np.abs(c - golden["cost"][:, 0]).max() <= 2e-14 * max(1.0, np.abs(golden["cost"]).max())
    assert np.abs(g - golden["cost_gradient"]).max() <= 2e-14
    assert np.abs(H - golden["cost_hessian"]).max() <= 2e-14
    assert abs(c[0] - 8.0) < 1e-14        # fcost(1...1) = 8
    c2 = robot5x2.cost(x, d)
    c3, g3 = robot5x2.cost_gradient(x, d)
    assert np.array_equal(c2, c) and np.array_equal(c3, c) and np.array_equal(g3, g)


def test_casadi_constraints_jacobian(robot5x2, golden):
    x = golden["x"]
    d = np.ones((x.shape[0], 1))
    c, J = robot5x2.equalities_linearised(x, d)
    assert np.abs(c - golden["constraints"]).max() <= 5e-14
    # the dense Jacobian's last block row is -reverse(first row) — equal to D's last row up to 2e-13
    assert np.abs(J - golden["constraints_jacobian"]).max() <= 5e-13
    assert np.array_equal(robot5x2.equalities(x, d), c)


def test_casadi_lagrangian_gradient_hessian(robot5x2, golden):
    x, lam_eq = golden["x"], golden["lam"]
    B = x.shape[0]
    d = np.ones((B, 1))
    rng = np.random.default_rng(5)
    lam = np.concatenate([lam_eq, rng.uniform(-1, 1, (B, 55))], axis=1)
    r = robot5x2.lagrangian_gradient_hessian(x, lam, d)
    # the CasADi Lagrangian omits the +lam_box term (codegen_test.cpp:260, 280 add it on the caller side)
    assert np.abs(r["lag_grad"] - lam[:, 33:] - golden["lagrangian_gradient"]).max() <= 5e-13
    assert np.abs(r["hess"] - golden["lagrangian_hessian"]).max() <= 5e-13
    r1 = robot5x2.lagrangian_gradient(x, lam, d)
    for k in ("cost", "lag_grad", "cost_grad", "g", "jac"):
        assert np.array_equal(r1[k], r[k]), k


# ---- QP known answers (reference tests/solvers/qp/box_admm_test.cpp) -------------------------------------------------------
def test_qp_simple(orc):
    st = orc.qp_default_settings(); st.max_iter = 150        # box_admm_test.cpp:15-45
    r = orc.qp_solve([[[4, 1], [1, 2]]], [[1, 1]], [[[1, 1]]], [[1]], [[1]], [[0, 0]], [[0.7, 0.7]], st)
    assert np.allclose(r["x"][0], [0.3, 0.7], rtol=1e-2)
    assert r["info"]["status"][0] == capi.QP_SOLVED and r["info"]["iter"][0] < 150


def test_qp_simple_lp(orc):
    st = orc.qp_default_settings()                          # box_admm_test.cpp:266-298
    st.max_iter = 200; st.alpha = 1.0; st.adaptive_rho = 1; st.check_termination = 10
    r = orc.qp_solve(np.zeros((1, 1, 1)), [[1.0]], np.zeros((1, 0, 1)), np.zeros((1, 0)), np.zeros((1, 0)), [[-1e6]], [[1e6]], st)
    assert np.allclose(r["x"][0], [-1e6], rtol=1e-2)
    assert r["info"]["status"][0] == capi.QP_SOLVED and r["info"]["iter"][0] < 200


def test_qp_nonconvex(orc):
    st = orc.qp_default_settings()                          # box_admm_test.cpp:300-334
    st.max_iter = 200; st.alpha = 1.0; st.adaptive_rho = 1; st.rho = 2; st.check_termination = 10
    r = orc.qp_solve(-np.ones((1, 1, 1)), [[0.0]], np.zeros((1, 0, 1)), np.zeros((1, 0)), np.zeros((1, 0)), [[-1.0]], [[2.0]], st,
                     x_guess=[[0.1]], y_guess=[[0.1]])
    assert np.allclose(r["x"][0], [2.0], rtol=1e-2)
    assert r["info"]["status"][0] == capi.QP_SOLVED and r["info"]["iter"][0] < 200


def test_constraint_classification(orc):
    """admm_solver_test.cpp:259-300 (parse_constraints_bounds, qp_base.hpp:195-222)"""
    st = orc.qp_default_settings(); st.max_iter = 25
    lo = [[-1e17, -101, -1e17, -1, 42]]
    hi = [[1e17, 1e17, 123, 1, 42]]
    r = orc.qp_solve(np.eye(5)[None], -np.ones((1, 5)), np.eye(5)[None], lo, hi, lo, hi, st)
    expect = [capi.LOOSE_BOUNDS, capi.INEQUALITY_CONSTRAINT, capi.INEQUALITY_CONSTRAINT, capi.INEQUALITY_CONSTRAINT, capi.EQUALITY_CONSTRAINT]
    assert list(r["ctype"][0][:5]) == expect and list(r["ctype"][0][5:]) == expect


def test_qp_matches_dense_kkt_solution(orc):
    """equality-constrained QP: ADMM fixed point == solution of the KKT system (independent numpy solve)"""
    rng = np.random.default_rng(11)
    N, M = 12, 5
    G = rng.standard_normal((N, N)); H = G @ G.T + N * np.eye(N)
    A = rng.standard_normal((M, N)); h = rng.standard_normal(N); b = rng.standard_normal(M)
    st = orc.qp_default_settings(); st.max_iter = 4000; st.eps_abs = 1e-9; st.eps_rel = 1e-9; st.check_termination = 10
    inf = np.full((1, N), np.inf)
    r = orc.qp_solve(H[None], h[None], A[None], b[None], b[None], -inf, inf, st)
    K = np.block([[H, A.T], [A, np.zeros((M, M))]])
    sol = np.linalg.solve(K, np.concatenate([-h, b]))
    assert r["info"]["status"][0] == capi.QP_SOLVED
    assert np.abs(r["x"][0] - sol[:N]).max() < 1e-6
    assert np.abs(r["y"][0][:M] - sol[N:]).max() < 1e-5


def test_ldlt_pivot_order_is_eigens_diagonal_rule(orc):
    """[Eigen-ext] Eigen 3.3.7 LDLT<Lower> (ldlt_inplace<Lower>::unblocked) is LEFT-looking: at step k it searches the largest
    |diagonal| of the trailing part BEFORE that part has received any update, swaps it to position k, and only then updates
    column k.  The pivot sequence is therefore a selection sort of |diag(K)| (first maximum wins).  Checked against a numpy
    replay of that rule, and the factorisation itself against numpy: P K P^T = L D L^T reproduces the ADMM iterate."""
    rng = np.random.default_rng(3)
    N, M = 7, 3
    G = rng.standard_normal((N, N)); H = G @ G.T + np.eye(N)
    A = rng.standard_normal((M, N))
    st = orc.qp_default_settings(); st.max_iter = 1; st.check_termination = 0; st.adaptive_rho = 0
    lo = np.full((1, N), -1.0); hi = np.full((1, N), 1.0)
    h = np.ones(N)
    r = orc.qp_solve(H[None], h[None], A[None], np.zeros((1, M)), np.zeros((1, M)), lo, hi, st)
    rho = 0.1
    n = N + M
    K = np.zeros((n, n))
    K[:N, :N] = H + (1e-6 + rho) * np.eye(N)
    K[N:, :N] = A; K[:N, N:] = A.T
    K[N:, N:] = -np.eye(M) / (1e3 * rho)
    dd = np.abs(np.diag(K)).copy()
    perm = np.arange(n)
    for k in range(n - 1):
        p = k + int(np.argmax(dd[k:]))
        dd[[k, p]] = dd[[p, k]]; perm[[k, p]] = perm[[p, k]]
    assert [int(v) for v in r["perm"][0]] == [int(v) for v in perm]
    # first ADMM trip from x = z = y = 0: K [x; nu] = [-h; 0]  (box_admm.hpp:351-355, 123-130)
    sol = np.linalg.solve(K, np.concatenate([-h, np.zeros(M)]))
    assert np.abs(r["x"][0] - sol[:N]).max() < 1e-12


# ---- BFGS (reference tests/solvers/sqp/bfgs_test.cpp:21-65) ----------------------------------------------------------------
@pytest.mark.parametrize("hdiag,converges", [((2.0, 1.0), True), ((2.0, -1.0), False)])
def test_bfgs_reference_cases(orc, hdiag, converges):
    H = np.diag(hdiag)
    B = np.eye(2)[None]
    for i in range(10):
        s = np.array([[np.sin(i), np.cos(i)]])
        y = s @ H
        B, branch = orc.bfgs_update(B, s, y)
        assert np.all(np.linalg.eigvalsh(B[0]) > 0)       # is_posdef
    if converges:
        assert np.allclose(B[0], H, rtol=1e-3, atol=1e-3)


def test_bfgs_formula(orc):
    rng = np.random.default_rng(2)
    n = 9
    G = rng.standard_normal((n, n)); B = G @ G.T + np.eye(n)
    s = rng.standard_normal(n); y = rng.standard_normal(n)
    for sign in (1.0, -1.0):    # plain and damped branches
        yy = sign * np.abs(s @ y) * y / (s @ y)
        Bn, br = orc.bfgs_update(B[None], s[None], yy[None])
        Bs = B @ s; sBs = s @ Bs; sy = s @ yy
        if sy < 0.2 * sBs:
            th = 0.8 * sBs / (sBs - sy); r = th * yy + (1 - th) * Bs; assert br[0] == 1
        else:
            r = yy; assert br[0] == 0
        ref = B - np.outer(Bs, Bs) / sBs + np.outer(r, r) / (s @ r)
        assert np.abs(Bn[0] - ref).max() < 1e-12 * np.abs(ref).max()


# ---- KKT assembly (box_admm.hpp:207-223) ------------------------------------------------------------------------------------
def test_kkt_assemble_layout(orc):
    rng = np.random.default_rng(4)
    B, N, M = 3, 6, 4
    H = rng.standard_normal((B, N, N)); A = rng.standard_normal((B, M, N))
    rb = rng.uniform(0.1, 1, (B, N)); ri = rng.uniform(0.1, 1, (B, M))
    K = orc.kkt_assemble(H, A, rb, ri, 1e-6)
    for b in range(B):
        K11 = H[b].copy()
        K11[np.arange(N), np.arange(N)] = (np.diag(H[b]) + 1e-6) + rb[b]
        assert np.array_equal(K[b, :N, :N], K11)
        assert np.array_equal(K[b, N:, :N], A[b])
        assert np.array_equal(K[b, :N, N:], np.zeros((N, M)))          # upper-right block is not written (Lower)
        assert np.array_equal(K[b, N:, N:], -np.diag(ri[b]))

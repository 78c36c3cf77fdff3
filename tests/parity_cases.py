"""Parity cases shared by the CPU suite (warp-emulator build of the kernels) and the GPU suite (the CUDA library).
Every case drives `api` and the oracle `orc` through the same C ABI on the same seeded inputs and demands BIT-EXACT
equality: integer decisions (pivot permutations, classifications, iteration counts, branches) and fp64 values alike —
both sides use the same deterministic elementary functions and the same summation orders (oracle/canon.hpp)."""
import numpy as np

from polympc_b200 import workloads as W

PROBLEM_SETUP = {
    "mobile_robot_6x2": dict(t=(0.0, 2.0), d=2.0),
    "mobile_robot_5x2": dict(t=(0.0, 1.0), d=1.0),
    "mobile_robot_5x3": dict(t=(0.0, 2.0), d=2.0),
    "cstr_5x2": dict(t=(0.0, 100.0), d=None),
    "kite_4x2": dict(t=(0.0, 0.5), d=4.0),
    "kite_12x1": dict(t=(0.0, 0.5), d=4.0),
    "robot_obstacle_5x2": dict(t=(0.0, 2.0), d=2.0),
    "parking_5x2": dict(t=(0.0, 1.0), d=1.0),
}


# reference-style problem classes compiled through include/polympc_compat/ (examples/dropin/*.hpp) and the hand-restated,
# fixture-pinned oracle model each of them must reproduce bit for bit
ORACLE_TWIN = {"dropin_robot_5x3": "mobile_robot_5x3", "dropin_cstr_5x2": "cstr_5x2"}


def sample_var(name, dims, B, rng):
    N, NX, NU, NN = dims["N"], dims["NX"], dims["NU"], dims["NN"]
    if name.startswith("mobile_robot") or name.startswith("robot_obstacle") or name.startswith("parking"):
        var = rng.uniform(-1.0, 1.0, (B, N))
        if name.startswith("parking"):
            var[:, -1] = rng.uniform(0.5, 3.0, B)          # the optimised final time
    elif name.startswith("cstr"):
        x = np.array([2.0, 1.0, 110.0, 108.0]) + rng.uniform(-1, 1, (B, NN, NX)) * np.array([0.5, 0.3, 5.0, 5.0])
        u = np.array([14.0, -1100.0]) + rng.uniform(-1, 1, (B, NN, NU)) * np.array([5.0, 500.0])
        var = np.concatenate([x.reshape(B, -1), u.reshape(B, -1)], axis=1)
    else:
        x = W.KITE_NOMINAL + rng.uniform(-1, 1, (B, NN, NX)) * 0.2
        u = np.array([1.5, 0.0, 0.0]) + rng.uniform(-1, 1, (B, NN, NU)) * 0.2
        var = np.concatenate([x.reshape(B, -1), u.reshape(B, -1)], axis=1)
    return var


def assert_same(a, b, what):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, what
    if a.dtype.kind == "f":
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        assert same.all(), f"{what}: {np.count_nonzero(~same)} of {a.size} entries differ, max abs diff {np.nanmax(np.abs(a - b))}"
    else:
        assert np.array_equal(a, b), f"{what}: integer arrays differ"


def ocp_case(api, orc, name, B=3, seed=0):
    """a3-a10: every transcription operator"""
    twin = ORACLE_TWIN.get(name, name)
    su = PROBLEM_SETUP[twin]
    rng = np.random.default_rng(seed)
    oa, ob = api.ocp(name), orc.ocp(twin)
    name = twin
    oa.set_time_limits(*su["t"]); ob.set_time_limits(*su["t"])
    assert_same(oa.time_nodes(), ob.time_nodes(), "time_nodes")
    D = oa.d
    var = sample_var(name, D, B, rng)
    d = None if su["d"] is None else np.full((B, D["ND"]), su["d"])
    lam = rng.uniform(-2, 2, (B, D["DUAL"]))
    assert_same(oa.cost(var, d), ob.cost(var, d), "cost")
    assert_same(oa.equalities(var, d), ob.equalities(var, d), "equalities")
    if D["NG"] > 0:
        assert_same(oa.inequalities(var, d), ob.inequalities(var, d), "inequalities")
    for x, y, n in zip(oa.equalities_linearised(var, d), ob.equalities_linearised(var, d), ("c", "jac")):
        assert_same(x, y, "equalities_linearised." + n)
    for x, y, n in zip(oa.cost_gradient(var, d), ob.cost_gradient(var, d), ("cost", "grad")):
        assert_same(x, y, "cost_gradient." + n)
    for x, y, n in zip(oa.cost_gradient_hessian(var, d), ob.cost_gradient_hessian(var, d), ("cost", "grad", "hess")):
        assert_same(x, y, "cost_gradient_hessian." + n)
    ra, rb = oa.lagrangian_gradient(var, lam, d), ob.lagrangian_gradient(var, lam, d)
    for k in rb:
        assert_same(ra[k], rb[k], "lagrangian_gradient." + k)
    ra, rb = oa.lagrangian_gradient_hessian(var, lam, d), ob.lagrangian_gradient_hessian(var, lam, d)
    for k in rb:
        assert_same(ra[k], rb[k], "lagrangian_gradient_hessian." + k)
    return rb


def random_qp(rng, B, N, M, n_eq=None, box=True, loose_rows=0):
    """strictly convex H, rows of A split into equalities / two-sided inequalities / loose rows; boxes partly infinite"""
    n_eq = M // 2 if n_eq is None else n_eq
    G = rng.standard_normal((B, N, N))
    H = G @ np.transpose(G, (0, 2, 1)) / N + np.eye(N)
    h = rng.standard_normal((B, N))
    A = rng.standard_normal((B, M, N))
    xf = rng.uniform(-0.5, 0.5, (B, N))
    Ax = np.einsum("bmn,bn->bm", A, xf)
    Alb = Ax - rng.uniform(0.1, 1.0, (B, M)); Aub = Ax + rng.uniform(0.1, 1.0, (B, M))
    Alb[:, :n_eq] = Ax[:, :n_eq]; Aub[:, :n_eq] = Ax[:, :n_eq]
    if loose_rows:
        Alb[:, M - loose_rows:] = -np.inf; Aub[:, M - loose_rows:] = np.inf
    if box:
        xlb = np.where(rng.uniform(size=(B, N)) < 0.5, -0.6, -np.inf); xub = np.where(rng.uniform(size=(B, N)) < 0.5, 0.6, np.inf)
    else:
        xlb = np.full((B, N), -np.inf); xub = np.full((B, N), np.inf)
    return H, h, A, Alb, Aub, xlb, xub


def qp_case(api, orc, N, M, B=3, seed=0, settings=None, warm=False, **kw):
    """a15-a21: boxADMM + pivoted LDLT; iterates, multipliers, active set, classification, pivot order, trip counts"""
    rng = np.random.default_rng(seed)
    H, h, A, Alb, Aub, xlb, xub = random_qp(rng, B, N, M, **kw)
    st = settings if settings is not None else orc.sqp_default_qp_settings()
    xg = rng.uniform(-0.1, 0.1, (B, N)) if warm else None
    yg = rng.uniform(-0.1, 0.1, (B, N + M)) if warm else None
    ra = api.qp_solve(H, h, A, Alb, Aub, xlb, xub, st, x_guess=xg, y_guess=yg)
    rb = orc.qp_solve(H, h, A, Alb, Aub, xlb, xub, st, x_guess=xg, y_guess=yg)
    for k in ("perm", "ctype", "n_factor"):
        assert_same(ra[k], rb[k], "qp." + k)
    for f in ("status", "iter", "rho_updates"):
        assert_same(ra["info"][f], rb["info"][f], "qp.info." + f)
    for f in ("rho_estimate", "res_prim", "res_dual"):
        assert_same(ra["info"][f], rb["info"][f], "qp.info." + f)
    for k in ("x", "y", "z", "q"):
        assert_same(ra[k], rb[k], "qp." + k)
    # active set: exact equality with the bounds on both sides
    act_a = np.concatenate([(ra["z"] == Alb) | (ra["z"] == Aub), (ra["q"] == xlb) | (ra["q"] == xub)], axis=1)
    act_b = np.concatenate([(rb["z"] == Alb) | (rb["z"] == Aub), (rb["q"] == xlb) | (rb["q"] == xub)], axis=1)
    assert np.array_equal(act_a, act_b)
    return rb


def admm_case(api, orc, N, M, B=3, seed=0, settings=None, warm=False, **kw):
    """the reference's OSQP-style ADMM<> (admm.hpp:112-213) as a batched operator: iterates, multipliers, the auxiliary vector z
    of size M + N and its active set, classification, pivot order of the (2N + M)-dimensional KKT system, trip counts"""
    rng = np.random.default_rng(seed)
    H, h, A, Alb, Aub, xlb, xub = random_qp(rng, B, N, M, **kw)
    st = settings if settings is not None else orc.sqp_default_qp_settings()
    xg = rng.uniform(-0.1, 0.1, (B, N)) if warm else None
    yg = rng.uniform(-0.1, 0.1, (B, N + M)) if warm else None
    ra = api.qp_solve_admm(H, h, A, Alb, Aub, xlb, xub, st, x_guess=xg, y_guess=yg)
    rb = orc.qp_solve_admm(H, h, A, Alb, Aub, xlb, xub, st, x_guess=xg, y_guess=yg)
    for k in ("perm", "ctype", "n_factor"):
        assert_same(ra[k], rb[k], "admm." + k)
    for f in ("status", "iter", "rho_updates", "rho_estimate", "res_prim", "res_dual"):
        assert_same(ra["info"][f], rb["info"][f], "admm.info." + f)
    for k in ("x", "y", "z"):
        assert_same(ra[k], rb[k], "admm." + k)
    lo, hi = np.concatenate([Alb, xlb], axis=1), np.concatenate([Aub, xub], axis=1)
    assert np.array_equal((ra["z"] == lo) | (ra["z"] == hi), (rb["z"] == lo) | (rb["z"] == hi))
    return rb


def bfgs_case(api, orc, N, B=4, seed=0):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((B, N, N))
    Bm = G @ np.transpose(G, (0, 2, 1)) / N + np.eye(N)
    s = rng.standard_normal((B, N)); y = rng.standard_normal((B, N))
    y[0] = np.einsum("ij,j->i", Bm[0], s[0])             # plain branch
    y[1] = -y[1]                                           # likely damped
    if B > 2:
        s[2] = 0.0                                         # skipped (sr < eps)
    Ba, bra = api.bfgs_update(Bm, s, y)
    Bb, brb = orc.bfgs_update(Bm, s, y)
    assert_same(bra, brb, "bfgs.branch")
    assert_same(Ba, Bb, "bfgs.B")
    return brb


def block_bfgs_case(api, orc, name, B=3, seed=0):
    """ContinuousOCP<..., SPARSE>::hessian_update_impl as an operator: H = exact Lagrangian Hessian of a random point (so it has
    the block pattern), random s; y chosen to hit the plain and the damped branch"""
    twin = ORACLE_TWIN.get(name, name)
    r = ocp_case(api, orc, name, B=B, seed=seed)
    oa, ob = api.ocp(name), orc.ocp(twin)
    N = oa.d["N"]
    rng = np.random.default_rng(seed + 100)
    H = r["hess"] + 2.0 * np.eye(N)
    s = rng.standard_normal((B, N)); y = rng.standard_normal((B, N))
    y[0] = np.einsum("ij,j->i", H[0], s[0])              # s'y = s'Hs: plain
    y[1] = -np.abs(y[1]) * np.sign(s[1])                 # s'y < 0: damped
    Ha, bra = oa.block_bfgs_update(H, s, y)
    Hb, brb = ob.block_bfgs_update(H, s, y)
    assert_same(bra, brb, "block_bfgs.branch")
    assert_same(Ha, Hb, "block_bfgs.H")
    return brb


def kkt_case(api, orc, N, M, B=3, seed=0):
    rng = np.random.default_rng(seed)
    H = rng.standard_normal((B, N, N)); A = rng.standard_normal((B, M, N))
    rb = rng.uniform(0.1, 1, (B, N)); ri = rng.uniform(0.1, 1, (B, M))
    assert_same(api.kkt_assemble(H, A, rb, ri, 1e-6), orc.kkt_assemble(H, A, rb, ri, 1e-6), "kkt")


def solve_workload(api, w, lo=0, hi=None, name=None, hessian_update=0, preconditioner=0, line_search=0, filter_beta=0.1, filter_depth=10,
                   solves=1, qp_solver=0):
    hi = w.batch if hi is None else hi
    s = api.sqp(name or w.name, hi - lo)
    W.configure(s, w, lo, hi)
    s.set_trace(True)
    if hessian_update:
        s.set_hessian_update(hessian_update)
    if preconditioner:
        s.set_preconditioner(preconditioner)
    if line_search:
        s.set_line_search(line_search, filter_beta, filter_depth)
    if qp_solver:
        s.set_qp_solver(qp_solver)
    for _ in range(solves):            # a second solve warm-starts from the kept iterate (and the kept filter)
        s.solve()
    out = dict(x=s.primal(), lam=s.dual(), info=s.info(), stats=s.stats(), trace=s.trace(w.sqp_max_iter),
               ms=s.last_solve_ms(), launches=s.last_solve_launches())
    if line_search:
        out["filter"] = s.filter()
    s.close()
    return out


def ruiz_case(api, orc, N, M, B=4, seed=0, variant=1):
    """RuizEquilibration::compute and both unscale overloads as operators (qp_preconditioners.hpp:151-300, 364-404): badly scaled
    data, an empty row / column (the zero guards), infinite box bounds"""
    rng = np.random.default_rng(seed)
    H = rng.standard_normal((B, N, N)); H = H + H.transpose(0, 2, 1)
    A = rng.standard_normal((B, M, N)) * 10.0 ** rng.integers(-3, 4, (B, M, 1))
    h = rng.standard_normal((B, N)) * 50
    if B > 1:
        H[1, :, N // 2] = 0; H[1, N // 2, :] = 0; A[1, :, N // 2] = 0      # a zero column of [H; A]
    if B > 2 and M > 0:
        A[2, M // 2, :] = 0                                              # a zero row of A
    if B > 3:
        H[3] *= 1e-6; A[3] *= 1e-5; h[3] = 0.0                           # below the SPARSE guard 1e-4, zero gradient
    l = -np.abs(rng.standard_normal((B, N))); u = np.abs(rng.standard_normal((B, N)))
    l[:, 0] = -np.inf; u[:, 1] = np.inf
    args = (H, h, A, -np.abs(rng.standard_normal((B, M))), np.abs(rng.standard_normal((B, M))), l, u)
    ra, rb = api.ruiz_equilibrate(variant, *args), orc.ruiz_equilibrate(variant, *args)
    for k in rb:
        assert_same(ra[k], rb[k], "ruiz." + k)
    x = rng.standard_normal((B, N)); y = rng.standard_normal((B, N + M))
    un = lambda a, r: a.ruiz_unscale(r["D"], r["E"], r["c"], r["H"], r["h"], r["A"], r["Al"], r["Au"], r["l"], r["u"], x=x, y=y)
    ua, ub = un(api, rb), un(orc, rb)
    for k in ub:
        assert_same(ua[k], ub[k], "ruiz_unscale." + k)
    # the round trip restores the QP data to rounding
    assert np.abs(ub["H"] - H).max() <= 1e-12 * max(1.0, np.abs(H).max()) and (M == 0 or np.abs(ub["A"] - A).max() <= 1e-12 * np.abs(A).max())
    return rb


def sqp_case(api, orc, w, hessian_update=0, **opts):
    """a11-a14, a22: whole SQP solves; iterates, multipliers, info and the per-iteration decision trace"""
    ra = solve_workload(api, w, hessian_update=hessian_update, **opts)
    rb = solve_workload(orc, w, name=ORACLE_TWIN.get(w.name, w.name), hessian_update=hessian_update, **opts)
    if "filter" in rb:
        assert_same(ra["filter"], rb["filter"], "sqp.filter")
    for f in ("iter", "qp_solver_iter", "status"):
        assert_same(ra["info"][f], rb["info"][f], "sqp.info." + f)
    for k in ("qp_iter", "bfgs", "ls_trials", "qp_factor", "alpha"):
        assert_same(ra["trace"][k], rb["trace"][k], "sqp.trace." + k)
    assert_same(ra["x"], rb["x"], "sqp.x")
    assert_same(ra["lam"], rb["lam"], "sqp.lam")
    assert_same(ra["stats"], rb["stats"], "sqp.stats")
    return ra, rb


def valet_parking_solve(api, x0_first, x0_second, preconditioner=2, line_search=1, name="mobile_robot_5x3", qp_solver=0):
    """the setup of reference tests/control/valet_parking_mpc_test.cpp:175-235 for a batch of (first, second) initial states:
    robot 5 x 3 on [0, 2], d = 2, SQP 10 / 10, QP max_iter 1000, filter beta 0.1, block BFGS (SPARSE problem), Ruiz equilibration
    (SPARSE), controls bounded on the last 11 nodes only (`tail(22)`), the state pinned at offset 30 (`segment(30, 3)`) — both as the
    test writes them —, then a second, warm-started solve from another pinned state.  Returns the results of both solves."""
    x0_first = np.atleast_2d(np.asarray(x0_first, dtype=np.float64)); x0_second = np.atleast_2d(np.asarray(x0_second, dtype=np.float64))
    B = x0_first.shape[0]
    s = api.sqp(name, B)
    d = s.d
    s.problem.set_time_limits(0.0, 2.0)
    st = s.settings(); st.max_iter = 10; st.line_search_max_iter = 10; s.set_settings(st)
    q = s.qp_settings(); q.max_iter = 1000; s.set_qp_settings(q)
    s.set_parameters(np.array([2.0]))
    s.set_trace(True)
    s.set_hessian_update(1); s.set_preconditioner(preconditioner); s.set_line_search(line_search, 0.1, 10)
    if qp_solver:
        s.set_qp_solver(qp_solver)
    N = d["N"]
    lb = np.full((B, N), -np.inf); ub = np.full((B, N), np.inf)
    ub[:, -22:] = np.tile([1.5, 0.75], 11); lb[:, -22:] = np.tile([-1.5, -0.75], 11)
    s.set_primal(np.zeros(N)); s.set_dual(np.zeros(d["DUAL"]))
    outs = []
    for x0 in (x0_first, x0_second):
        lb[:, 30:33] = x0; ub[:, 30:33] = x0
        s.set_bounds_x(lb, ub)
        s.solve()
        outs.append(dict(x=s.primal(), lam=s.dual(), info=s.info(), stats=s.stats(), trace=s.trace(10),
                         filter=s.filter() if line_search else None))
    s.close()
    return outs


def valet_parking_case(api, orc, x0_first, x0_second, **kw):
    ra, rb = valet_parking_solve(api, x0_first, x0_second, **kw), valet_parking_solve(orc, x0_first, x0_second, **kw)
    for k, (a, b) in enumerate(zip(ra, rb)):
        for f in ("iter", "qp_solver_iter", "status"):
            assert_same(a["info"][f], b["info"][f], f"valet solve {k}: info." + f)
        for f in ("qp_iter", "bfgs", "ls_trials", "qp_factor", "alpha"):
            assert_same(a["trace"][f], b["trace"][f], f"valet solve {k}: trace." + f)
        assert_same(a["x"], b["x"], f"valet solve {k}: x"); assert_same(a["lam"], b["lam"], f"valet solve {k}: lam")
        if b["filter"] is not None:
            assert_same(a["filter"], b["filter"], f"valet solve {k}: filter")
    return ra, rb

"""GPU suite, fast arithmetic (PMB_ARITH_FAST: tile LDL^T on the fp64 tensor cores, explicit inverse of the unit-lower factor,
csrc/pmb_qp_fast.hpp) against the exact path and the CPU oracle.

What can be demanded of a path that rounds differently (SURVEY.md §8d "parity definitions"):
  * discrete decisions of ONE stage from identical inputs are identical: pivot permutation, bound classification, ADMM trip
    count, factorisation count — asserted on every instance;
  * stage-wise iterate parity: after one QP / one SQP iteration from the same state, max rel-inf <= 1e-10 (the north_star
    tolerance, TOL below) — asserted on EVERY instance of the full-size batches;
  * whole solves: rounding differences are amplified by the SQP iteration (the last iterations run Armijo tests and
    termination tests at round-off level), so per-iteration decision traces agree for only ~40 % of the instances under ANY
    change of rounding — the CPU oracle compiled with other flags shows the same numbers against itself
    (tests/test_oracle_flavours.py).  What is asserted for whole solves: identical status on >= 99.9 %, identical iteration
    count on >= 99.9 %, the median instance within TOL, and the solution quality (cost, constraint violation) unchanged."""
import numpy as np
import pytest

import parity_cases as pc
from polympc_b200 import workloads as W

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel(a, b):
    return np.max(np.abs(a - b), axis=1) / np.maximum(1.0, np.max(np.abs(b), axis=1))


@pytest.fixture()
def fast_default(pmb):
    pmb.set_default_arithmetic(1)
    yield pmb
    pmb.set_default_arithmetic(0)


@pytest.mark.parametrize("N,M", [(1, 0), (2, 1), (5, 5), (12, 7), (33, 20), (40, 31), (65, 39), (66, 44), (100, 92)])
def test_fast_qp_vs_oracle(fast_default, orc, N, M):
    """boxADMM with the fast linear algebra: every discrete decision identical to the oracle, iterates to rounding"""
    rng = np.random.default_rng(N)
    B = 33
    H, h, A, Alb, Aub, xlb, xub = pc.random_qp(rng, B, N, M)
    st = orc.sqp_default_qp_settings()
    ra = fast_default.qp_solve(H, h, A, Alb, Aub, xlb, xub, st)
    rb = orc.qp_solve(H, h, A, Alb, Aub, xlb, xub, st)
    for k in ("perm", "ctype", "n_factor"):
        pc.assert_same(ra[k], rb[k], "qp." + k)
    for f in ("status", "iter", "rho_updates"):
        pc.assert_same(ra["info"][f], rb["info"][f], "qp.info." + f)
    # iterates after up to 100 ADMM trips of a RANDOM dense QP (condition numbers up to ~1e6): rounding level, 1e-8; the 1e-10
    # bar is asserted below on the benchmark problems, per SQP iteration
    for k, tol in (("x", 1e-8), ("z", 1e-8), ("q", 1e-8), ("y", 1e-6)):       # multipliers of near-degenerate rows are the loosest
        if ra[k].size:
            assert rel(ra[k].reshape(B, -1), rb[k].reshape(B, -1)).max() <= tol, k
    # active set: exact equality with the bounds on both sides
    act_a = np.concatenate([(ra["z"] == Alb) | (ra["z"] == Aub), (ra["q"] == xlb) | (ra["q"] == xub)], axis=1)
    act_b = np.concatenate([(rb["z"] == Alb) | (rb["z"] == Aub), (rb["q"] == xlb) | (rb["q"] == xub)], axis=1)
    assert np.array_equal(act_a, act_b)


def _solve(api, w, arithmetic, trace=False, hessian_update=0, lo=0, hi=None, preconditioner=0, line_search=0):
    hi = w.batch if hi is None else hi
    s = api.sqp(w.name, hi - lo)
    W.configure(s, w, lo, hi)
    if arithmetic:
        s.set_arithmetic(arithmetic)
    if hessian_update:
        s.set_hessian_update(hessian_update)
    if preconditioner:
        s.set_preconditioner(preconditioner)
    if line_search:
        s.set_line_search(line_search, 0.1, 10)
    if trace:
        s.set_trace(True)
    s.solve()
    out = dict(x=s.primal(), lam=s.dual(), info=s.info(), stats=s.stats())
    if trace:
        out["trace"] = s.trace(w.sqp_max_iter)
    s.close()
    return out


@pytest.mark.parametrize("workload,batch", [("mobile_robot", 8192), ("cstr", 4096)])
def test_fast_one_sqp_iteration_every_instance(pmb, orc, workload, batch):
    """stage-wise parity at BASELINE.json's full sizes: one SQP iteration (exact Hessian, QP, line search, step) from the same
    state — every instance within 1e-10 rel-inf of the exact path, identical ADMM trip counts and line-search decisions; a
    256-instance sample of the exact path is in turn bit-identical to the oracle"""
    w = W.WORKLOADS[workload](batch, sqp_max_iter=1)
    a, b = _solve(pmb, w, 0, trace=True), _solve(pmb, w, 1, trace=True)
    for k in ("qp_iter", "ls_trials", "qp_factor"):
        pc.assert_same(a["trace"][k], b["trace"][k], k)
    pc.assert_same(a["trace"]["alpha"], b["trace"]["alpha"], "alpha")
    e = np.maximum(rel(b["x"], a["x"]), rel(b["lam"], a["lam"]))
    print(f"{workload}: one SQP iteration, fast vs exact: max rel-inf {e.max():.3e} over {batch} instances")
    assert e.max() <= TOL
    o = _solve(orc, w, 0, lo=0, hi=256)
    pc.assert_same(a["x"][:256], o["x"], "exact vs oracle")


def _trace_identical(a, b):
    same = (a["info"]["iter"] == b["info"]["iter"]) & (a["info"]["status"] == b["info"]["status"])
    for k in ("qp_iter", "bfgs", "ls_trials", "qp_factor"):
        same &= (a["trace"][k] == b["trace"][k]).all(axis=1)
    same &= np.array([np.array_equal(u, v, equal_nan=True) for u, v in zip(a["trace"]["alpha"], b["trace"]["alpha"])])
    return same


@pytest.mark.parametrize("workload,batch", [("mobile_robot", 8192), ("cstr", 4096)])
def test_fast_whole_solves_against_the_oracle(pmb, orc, workload, batch):
    """ALL instances of BASELINE.json configs 2 and 3, fast arithmetic on the GPU vs the CPU oracle (every straggler included)"""
    from oracle import pyoracle
    import os
    pyoracle.set_num_threads(os.cpu_count() or 1)
    w = W.WORKLOADS[workload](batch)
    f, o = _solve(pmb, w, 1, trace=True), _solve(orc, w, 0, trace=True)
    pyoracle.set_num_threads(1)
    same_status = (f["info"]["status"] == o["info"]["status"]).mean()
    same_iter = (f["info"]["iter"] == o["info"]["iter"]).mean()
    ident = _trace_identical(f, o)
    fin = np.isfinite(o["x"]).all(axis=1) & np.isfinite(f["x"]).all(axis=1)
    e = np.maximum(rel(f["x"], o["x"]), rel(f["lam"], o["lam"]))
    m = ident & fin
    print(f"{workload} x {batch}, fast GPU vs oracle: identical status {100 * same_status:.3f} %, identical iteration count "
          f"{100 * same_iter:.3f} %, identical decision traces {100 * ident.mean():.2f} %; on those: max rel-inf {e[m].max():.3e}, "
          f"median {np.median(e[m]):.3e}, within 1e-10: {100 * (e[m] <= TOL).mean():.2f} %; all finite instances: max {e[fin].max():.3e}")
    assert same_status >= 0.999 and same_iter >= 0.999
    assert np.median(e[m]) <= TOL and (e[m] <= TOL).mean() >= 0.90
    # solution quality is unchanged: final cost and constraint violation of the solved instances
    solved = (o["info"]["status"] == 0) & (f["info"]["status"] == 0)
    dc = np.abs(f["stats"][solved, 0] - o["stats"][solved, 0]) / (1.0 + np.abs(o["stats"][solved, 0]))
    assert np.percentile(dc, 99.9) <= 1e-6
    assert f["stats"][solved, 3].max() <= 1e-3 + 1e-12          # max constraint violation <= eps_prim on both sides
    assert np.isfinite(f["x"]).all(axis=1).sum() == np.isfinite(o["x"]).all(axis=1).sum()


def test_fast_block_bfgs_and_minimal_time(pmb, orc):
    """the other Hessian modes run on the fast path too: block BFGS (robot) and exact Hessian + Gershgorin with NP = 1 (parking)"""
    w = W.mobile_robot(512, sqp_max_iter=1)
    a, b = _solve(pmb, w, 0, hessian_update=1), _solve(pmb, w, 1, hessian_update=1)
    assert np.maximum(rel(b["x"], a["x"]), rel(b["lam"], a["lam"])).max() <= TOL
    w = W.parking(256)
    a, b = _solve(pmb, w, 0), _solve(pmb, w, 1)
    assert (a["info"]["status"] == b["info"]["status"]).mean() >= 0.99
    assert (b["info"]["status"] == 0).mean() >= 0.85


def test_fast_ruiz_and_filter_line_search(pmb, orc):
    """Ruiz equilibration and the filter line search on the fast path: one SQP iteration agrees with the exact path to tolerance
    (same ADMM trip counts and step lengths), whole solves converge like the oracle's"""
    w = W.mobile_robot(1024, sqp_max_iter=1)
    a, b = _solve(pmb, w, 0, trace=True, preconditioner=2, line_search=1), _solve(pmb, w, 1, trace=True, preconditioner=2, line_search=1)
    pc.assert_same(a["trace"]["qp_iter"], b["trace"]["qp_iter"], "qp_iter"); pc.assert_same(a["trace"]["alpha"], b["trace"]["alpha"], "alpha")
    assert np.maximum(rel(b["x"], a["x"]), rel(b["lam"], a["lam"])).max() <= TOL
    w = W.mobile_robot(1024, sqp_max_iter=20, ls_max_iter=20)
    f, o = _solve(pmb, w, 1, preconditioner=1, line_search=1), _solve(orc, w, 0, preconditioner=1, line_search=1)
    assert (f["info"]["status"] == o["info"]["status"]).mean() >= 0.99 and (f["info"]["status"] == 0).mean() > 0.9


def test_fast_kite_is_available_but_not_within_tolerance(pmb):
    """kite 12 x 1 (K = 377; tile workspace in a global, L2-resident slot): the explicit inverse of a 377 x 377 unit-lower factor
    loses more than the tolerance allows — measured 2e-7 per SQP iteration.  The engine runs it (2x faster), but bench.py and
    the documentation quote the kite in exact arithmetic.  Pinned here so that the statement stays true."""
    w = W.kite(64, sqp_max_iter=1)
    a, b = _solve(pmb, w, 0), _solve(pmb, w, 1)
    e = np.maximum(rel(b["x"], a["x"]), rel(b["lam"], a["lam"])).max()
    print(f"kite: one SQP iteration, fast vs exact: max rel-inf {e:.3e}")
    assert TOL < e <= 1e-4
    pc.assert_same(a["info"]["qp_solver_iter"], b["info"]["qp_solver_iter"], "ADMM trip counts")

"""Synthetic workloads of BASELINE.json: problem data and seeded initial-state sweeps.

Host-side product code (numpy only).  Each workload returns everything a batched solver needs; `configure()` applies it to
any object with the Sqp interface of capi.py (the CUDA engine — or, in tests and the CPU-baseline leg, the oracle).
Reference setups: tests/control/mpc_wrapper_test.cpp:120-140 (robot), tests/control/cstr_control_test.cpp:137-177 (CSTR);
the kite is our own model (the reference ships none).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Workload:
    name: str                 # registry name of the problem (pmb_problem_name)
    t0: float
    tf: float
    d: np.ndarray | None      # static parameters (ND,) or None
    u_lb: np.ndarray
    u_ub: np.ndarray
    x0: np.ndarray            # (batch, NX) initial states
    x_guess: np.ndarray | None = None    # (NX,) constant state guess (None: zeros, like the reference SQPBase ctor)
    u_guess: np.ndarray | None = None
    sqp_max_iter: int = 100
    ls_max_iter: int = 100
    g_lb: np.ndarray | None = None       # (NG,) bounds of the generic inequality constraints, same at every node
    g_ub: np.ndarray | None = None
    p_lb: np.ndarray | None = None       # (NP,) bounds / guess of the optimised parameters (MPC::parameters_bounds, p_guess)
    p_ub: np.ndarray | None = None
    p_guess: np.ndarray | None = None
    xf_lb: np.ndarray | None = None      # (NX,) bounds of the FINAL state (MPC::final_state_bounds: node 0 of the X block)
    xf_ub: np.ndarray | None = None
    exact_hessian: bool = False          # SQPBase overrides of the reference's minimal_time_test.cpp:90-143
    gershgorin: bool = False
    meta: dict = field(default_factory=dict)

    @property
    def batch(self) -> int:
        return self.x0.shape[0]


def _blocks(batch: int, block: int, seed: int, draw) -> np.ndarray:
    """`batch` rows drawn in blocks of `block` rows, block k from its own stream (seed + 1000 k): the first rows of a sweep are the
    same whatever the total size, so that N GPUs x `block` instances each is N times the single-GPU workload plus new instances
    (rank r of a weak-scaling run owns block r) and a 1 -> 8 curve measures scaling, not a different draw of stragglers."""
    rows = []
    k = 0
    while sum(len(r) for r in rows) < batch:
        rows.append(draw(np.random.default_rng(seed + 1000 * k), block))
        k += 1
    return np.ascontiguousarray(np.concatenate(rows, axis=0)[:batch])


def mobile_robot(batch: int, seed: int = 20260117 + 2, grid: str = "6x2", sqp_max_iter: int = 100, ls_max_iter: int = 100) -> Workload:
    x0 = _blocks(batch, 8192, seed, lambda rng, n: np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n),
                                                                   rng.uniform(-np.pi / 4, np.pi / 4, n)]))
    return Workload(f"mobile_robot_{grid}", 0.0, 2.0, np.array([2.0]), np.array([-1.5, -0.75]), np.array([1.5, 0.75]), x0,
                    sqp_max_iter=sqp_max_iter, ls_max_iter=ls_max_iter,
                    meta={"x0": "U([-1,1]^2 x [-pi/4,pi/4])", "seed": seed})


def robot_obstacle(batch: int, seed: int = 20260117 + 6, sqp_max_iter: int = 100, ls_max_iter: int = 100) -> Workload:
    """mobile robot 5x2 that has to keep a distance of 0.3 from an obstacle at (0.25, 0.25): NG = 1 (not a reference case)"""
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0, 2 * np.pi, batch)
    rad = rng.uniform(0.6, 1.0, batch)
    x0 = np.column_stack([0.25 + rad * np.cos(ang), 0.25 + rad * np.sin(ang), rng.uniform(-np.pi / 4, np.pi / 4, batch)])
    return Workload("robot_obstacle_5x2", 0.0, 2.0, np.array([2.0]), np.array([-1.5, -0.75]), np.array([1.5, 0.75]), x0,
                    sqp_max_iter=sqp_max_iter, ls_max_iter=ls_max_iter, g_lb=np.array([0.09]), g_ub=np.array([np.inf]),
                    meta={"x0": "ring of radius 0.6-1.0 around the obstacle", "seed": seed})


def cstr(batch: int, seed: int = 20260117 + 3, sqp_max_iter: int = 100, ls_max_iter: int = 100) -> Workload:
    nominal = np.array([1.0, 0.5, 100.0, 100.0])
    spread = np.array([0.1, 0.05, 1.0, 1.0])
    x0 = _blocks(batch, 4096, seed, lambda rng, n: nominal + rng.uniform(-1, 1, (n, 4)) * spread)
    return Workload("cstr_5x2", 0.0, 100.0, None, np.array([3.0, -9000.0]), np.array([35.0, 0.0]), x0,
                    sqp_max_iter=sqp_max_iter, ls_max_iter=ls_max_iter,
                    meta={"x0": "(1,0.5,100,100) + U(+-(0.1,0.05,1,1))", "seed": seed})


KITE_NOMINAL = np.array([12.0, 0.0, 0.5, 0.0, 0.0, 0.0, 0.0, 0.0, -50.0, 1.0, 0.0, 0.0, 0.0])


def kite(batch: int, seed: int = 20260117 + 4, grid: str = "12x1", sqp_max_iter: int = 100, ls_max_iter: int = 100) -> Workload:
    def draw(rng, n):
        x = KITE_NOMINAL * (1.0 + 0.05 * rng.uniform(-1, 1, (n, 13)))
        x[:, [1, 3, 4, 5, 6, 7, 10, 11, 12]] += 0.05 * rng.uniform(-1, 1, (n, 9))   # entries whose nominal value is 0
        return x
    x0 = _blocks(batch, 1024, seed, draw)
    # horizon 0.5 s: with the default SQP/QP settings every instance of the sweep converges (4-13 SQP iterations)
    return Workload(f"kite_{grid}", 0.0, 0.5, np.array([4.0]), np.array([0.0, -0.3, -0.3]), np.array([5.0, 0.3, 0.3]), x0,
                    x_guess=KITE_NOMINAL, u_guess=np.array([1.5, 0.0, 0.0]),
                    sqp_max_iter=sqp_max_iter, ls_max_iter=ls_max_iter,
                    meta={"x0": "nominal +- 5 %", "seed": seed})


def parking(batch: int, seed: int = 20260117 + 7, sqp_max_iter: int = 20, ls_max_iter: int = 10) -> Workload:
    """free-final-time valet parking of the reference's tests/control/minimal_time_test.cpp:146-188 (NP = 1): drive from x0 into a
    +-0.05 box around the origin in minimal time; horizon normalised to [0, 1], the final time is the optimised parameter.
    Instance 0 is the reference's own x0 = (1.5, 0.5, 0.5); the others are drawn around it."""
    rng = np.random.default_rng(seed)
    x0 = np.array([1.5, 0.5, 0.5]) + rng.uniform(-1, 1, (batch, 3)) * np.array([0.3, 0.3, 0.2])
    x0[0] = [1.5, 0.5, 0.5]
    return Workload("parking_5x2", 0.0, 1.0, np.array([1.0]), np.array([-1.5, -0.75]), np.array([1.5, 0.75]), x0,
                    sqp_max_iter=sqp_max_iter, ls_max_iter=ls_max_iter, p_lb=np.array([0.0]), p_ub=np.array([10.0]), p_guess=np.array([0.5]),
                    xf_lb=np.full(3, -0.05), xf_ub=np.full(3, 0.05), exact_hessian=True, gershgorin=True,
                    meta={"x0": "(1.5,0.5,0.5) + U(+-(0.3,0.3,0.2))", "x_guess": "x0 at every node (minimal_time_test.cpp:166)", "seed": seed})


WORKLOADS = {"mobile_robot": mobile_robot, "cstr": cstr, "kite": kite, "robot_obstacle": robot_obstacle, "parking": parking}


def bounds_x(dims: dict, w: Workload):
    """Box bounds of the NLP variable [X | U]: states free, controls boxed (MPC::control_bounds, mpc_wrapper.hpp:120-135)."""
    N, NX, NU, NN = dims["N"], dims["NX"], dims["NU"], dims["NN"]
    lbx = np.full(N, -np.inf)
    ubx = np.full(N, np.inf)
    lbx[NX * NN:NX * NN + NU * NN] = np.tile(w.u_lb, NN)
    ubx[NX * NN:NX * NN + NU * NN] = np.tile(w.u_ub, NN)
    if dims["NP"] > 0 and w.p_lb is not None:
        lbx[NX * NN + NU * NN:] = w.p_lb
        ubx[NX * NN + NU * NN:] = w.p_ub
    if w.xf_lb is not None:                                  # MPC::final_state_bounds (mpc_wrapper.hpp): node 0 = final time
        lbx[:NX] = w.xf_lb
        ubx[:NX] = w.xf_ub
    return lbx, ubx


def configure(solver, w: Workload, lo: int = 0, hi: int | None = None) -> None:
    """Apply workload rows [lo, hi) to `solver` (an Sqp of capi.py with batch == hi - lo)."""
    hi = w.batch if hi is None else hi
    d = solver.d
    solver.problem.set_time_limits(w.t0, w.tf)
    st = solver.settings()
    st.max_iter = w.sqp_max_iter
    st.line_search_max_iter = w.ls_max_iter
    solver.set_settings(st)
    lbx, ubx = bounds_x(d, w)
    solver.set_bounds_x(lbx, ubx)
    if d["NG"] > 0:
        ng = d["NG"] * d["NN"]
        lbg = np.full(ng, -np.inf) if w.g_lb is None else np.tile(np.asarray(w.g_lb, dtype=np.float64), d["NN"])
        ubg = np.full(ng, np.inf) if w.g_ub is None else np.tile(np.asarray(w.g_ub, dtype=np.float64), d["NN"])
        solver.set_bounds_g(lbg, ubg)
    if d["ND"] > 0:
        solver.set_parameters(np.asarray(w.d, dtype=np.float64))
    guess = np.zeros(d["N"])
    if w.x_guess is not None:
        guess[:d["NX"] * d["NN"]] = np.tile(w.x_guess, d["NN"])
    if w.u_guess is not None:
        guess[d["NX"] * d["NN"]:d["NX"] * d["NN"] + d["NU"] * d["NN"]] = np.tile(w.u_guess, d["NN"])
    if w.exact_hessian or w.gershgorin:
        solver.set_hessian_options(w.exact_hessian, w.gershgorin)
    if w.name.startswith("parking"):
        # minimal_time_test.cpp:165-166: p_guess(0.5), x_guess(x0.replicate(11, 1)) — one guess per instance
        g = np.tile(guess, (hi - lo, 1))
        g[:, :d["NX"] * d["NN"]] = np.tile(w.x0[lo:hi], (1, d["NN"]))
        g[:, d["NX"] * d["NN"] + d["NU"] * d["NN"]:] = w.p_guess
        solver.set_primal(g)
    else:
        solver.set_primal(guess)
    solver.set_dual(np.zeros(d["DUAL"]))
    solver.set_initial_conditions(w.x0[lo:hi])


def shard_bounds(batch: int, world_size: int, rank: int):
    """Contiguous block of ceil(batch / world_size) instances per rank; the last ranks may get less."""
    per = -(-batch // world_size)
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)

"""tools/fast_check.py — run one workload in exact and fast arithmetic on the GPU: device time per batch, in-kernel phase cycles,
and how the two sets of results relate (to each other and, on a sample, to the CPU oracle).  Development / profiling aid."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import polympc_b200  # noqa: E402
from polympc_b200 import workloads as W  # noqa: E402


def rel(a, b):
    return np.max(np.abs(a - b), axis=1) / np.maximum(1.0, np.max(np.abs(b), axis=1))


def run(api, w, mode, profile, reps=3, trace=False, hessian_update=0):
    s = api.sqp(w.name, w.batch)
    W.configure(s, w)
    s.set_arithmetic(mode)
    if hessian_update:
        s.set_hessian_update(hessian_update)
    if trace:
        s.set_trace(True)
    ms = []
    for _ in range(reps):
        s.reset_guess(); s.solve(); ms.append(s.last_solve_ms())
    out = dict(x=s.primal(), lam=s.dual(), info=s.info(), ms=min(ms))
    if trace:
        out["trace"] = s.trace(w.sqp_max_iter)
    if profile:
        s.set_profiling(True); s.reset_guess(); s.solve(); out["cycles"] = s.phase_cycles()
    s.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mobile_robot")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--oracle", type=int, default=512, help="instances also solved by the CPU oracle")
    ap.add_argument("--max-iter", type=int, default=100)
    ap.add_argument("--block-bfgs", type=int, default=0)
    args = ap.parse_args()
    api = polympc_b200.load()
    w = W.WORKLOADS[args.workload](args.batch, sqp_max_iter=args.max_iter) if args.workload != "parking" else W.parking(args.batch)
    res = {}
    for name, mode in (("exact", 0), ("fast", 1)):
        r = run(api, w, mode, True, trace=True, hessian_update=args.block_bfgs)
        it = int(r["info"]["iter"].sum())
        res[name] = r
        cyc = r["cycles"]; n = max(1, cyc["sqp_iterations"])
        per_it = {k: round(v / n) for k, v in cyc.items() if k not in ("sqp_iterations", "admm_trips")}
        print(json.dumps({"mode": name, "ms_per_batch": round(r["ms"], 3), "sqp_iterations": it, "it_per_s": round(it / (r["ms"] * 1e-3)),
                          "solved": float((r["info"]["status"] == 0).mean()), "cycles_per_sqp_iteration": per_it,
                          "admm_trips_per_iteration": round(cyc["admm_trips"] / n, 2)}))
    a, b = res["exact"], res["fast"]
    same = (a["info"]["iter"] == b["info"]["iter"]) & (a["info"]["status"] == b["info"]["status"])
    for k in ("qp_iter", "bfgs", "ls_trials", "qp_factor"):
        same &= (a["trace"][k] == b["trace"][k]).all(1)
    same &= np.array([np.array_equal(u, v, equal_nan=True) for u, v in zip(a["trace"]["alpha"], b["trace"]["alpha"])])
    fin = np.isfinite(a["x"]).all(1) & np.isfinite(b["x"]).all(1)
    e = np.maximum(rel(b["x"], a["x"]), rel(b["lam"], a["lam"]))
    m = same & fin
    print(json.dumps({"fast_vs_exact": {"identical_decision_traces_pct": round(100 * same.mean(), 3), "identical_status_pct": round(100 * (a["info"]["status"] == b["info"]["status"]).mean(), 3),
                                        "identical_iter_pct": round(100 * (a["info"]["iter"] == b["info"]["iter"]).mean(), 3),
                                        "max_rel_inf_on_identical_traces": float(e[m].max()) if m.any() else None,
                                        "p50": float(np.median(e[m])) if m.any() else None, "p99": float(np.percentile(e[m], 99)) if m.any() else None,
                                        "within_1e-10_pct_of_identical": round(100 * float((e[m] <= 1e-10).mean()), 3) if m.any() else None,
                                        "max_rel_inf_all_finite": float(e[fin].max())}}))
    if args.oracle > 0:
        from oracle import pyoracle
        orc = pyoracle.load(); pyoracle.set_num_threads(os.cpu_count() or 1)
        nb = min(args.oracle, args.batch)
        s = orc.sqp(w.name, nb); W.configure(s, w, 0, nb)
        if args.block_bfgs:
            s.set_hessian_update(args.block_bfgs)
        t = time.time(); s.solve(); dt = time.time() - t
        xo, lo, io = s.primal(), s.dual(), s.info(); s.close()
        print(json.dumps({"oracle": {"instances": nb, "seconds": round(dt, 2), "it_per_s": round(int(io["iter"].sum()) / dt),
                                     "exact_bit_identical_pct": round(100 * float(((a["x"][:nb] == xo) | (np.isnan(a["x"][:nb]) & np.isnan(xo))).all(1).mean()), 3),
                                     "fast_max_rel_inf": float(np.nanmax(np.maximum(rel(b["x"][:nb], xo), rel(b["lam"][:nb], lo)))),
                                     "fast_identical_iter_pct": round(100 * float((b["info"]["iter"][:nb] == io["iter"]).mean()), 3)}}))
    # one SQP iteration from the same state: the stage-wise parity of SURVEY.md §8d
    w1 = W.WORKLOADS[args.workload](min(args.batch, 2048), sqp_max_iter=1) if args.workload != "parking" else None
    if w1 is not None:
        a1, b1 = run(api, w1, 0, False, reps=1), run(api, w1, 1, False, reps=1)
        e1 = np.maximum(rel(b1["x"], a1["x"]), rel(b1["lam"], a1["lam"]))
        print(json.dumps({"one_sqp_iteration_fast_vs_exact": {"instances": w1.batch, "max_rel_inf": float(e1.max()), "p99": float(np.percentile(e1, 99)),
                                                            "identical_qp_iterations_pct": round(100 * float((a1["info"]["qp_solver_iter"] == b1["info"]["qp_solver_iter"]).mean()), 3)}}))


if __name__ == "__main__":
    main()

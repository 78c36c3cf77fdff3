"""One batched solve for profilers (ncu): python tools/profile_step.py [workload] [batch] [solves]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polympc_b200  # noqa: E402
from polympc_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mobile_robot"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
solves = int(sys.argv[3]) if len(sys.argv) > 3 else 1
api = polympc_b200.load()
w = W.WORKLOADS[name](batch)
s = api.sqp(w.name, batch)
W.configure(s, w)
import os as _os
max_iter = int(_os.environ.get("PMB_MAX_ITER", "100"))
if max_iter != 100:
    st = s.settings(); st.max_iter = max_iter; s.set_settings(st)
s.set_arithmetic(int(_os.environ.get("PMB_ARITH", "0")))          # 0 exact, 1 fast
if int(_os.environ.get("PMB_BLOCK_BFGS", "0")):
    s.set_hessian_update(1)
if int(_os.environ.get("PMB_PRECOND", "0")):
    s.set_preconditioner(int(_os.environ["PMB_PRECOND"]))        # 1 Ruiz dense, 2 Ruiz sparse
if int(_os.environ.get("PMB_FILTER_LS", "0")):
    s.set_line_search(1, 0.1, 4)
s.set_profiling(bool(int(_os.environ.get("PMB_PROFILE", "1"))))
for _ in range(solves):
    s.reset_guess()
    s.solve()
ph = s.phase_cycles()
it = max(1, ph["sqp_iterations"])
print("cycles per SQP iteration:", {k: round(v / it) for k, v in ph.items() if k not in ("sqp_iterations", "admm_trips")},
      "ADMM trips/iteration %.1f" % (ph["admm_trips"] / it))
info = s.info()
print(f"{w.name} batch={batch}: {int(info['iter'].sum())} SQP iterations, {s.last_solve_ms():.2f} ms, {s.last_solve_launches()} launches, "
      f"solved {float((info['status'] == 0).mean()):.4f}")

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
for _ in range(solves):
    s.reset_guess()
    s.solve()
info = s.info()
print(f"{w.name} batch={batch}: {int(info['iter'].sum())} SQP iterations, {s.last_solve_ms():.2f} ms, {s.last_solve_launches()} launches, "
      f"solved {float((info['status'] == 0).mean()):.4f}")

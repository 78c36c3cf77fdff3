import sys; sys.path.insert(0,'/root/repo')
import polympc_b200
from polympc_b200 import workloads as W
api=polympc_b200.load()
w=W.mobile_robot(8192)
s=api.sqp(w.name,w.batch); W.configure(s,w); s.set_arithmetic(int(sys.argv[1]) if len(sys.argv)>1 else 1)
for _ in range(2): s.reset_guess(); s.solve()
print(s.last_solve_ms(), s.info()['iter'].sum())

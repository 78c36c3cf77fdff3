import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polympc_b200
from polympc_b200 import workloads as W
from oracle import pyoracle
pmb = polympc_b200.load(); orc = pyoracle.load()
for grid in ("5x3", "6x2", "5x2"):
    w = W.mobile_robot(1, grid=grid, sqp_max_iter=10, ls_max_iter=10)
    w.x0[:] = [0.5, 0.5, 0.5]
    outs = []
    for api in (pmb, orc):
        s = api.sqp(w.name, 1)
        s.problem.set_params(np.array([2, 2, 2, 1, 1, 1, 1, 1.0]))
        W.configure(s, w); s.solve()
        outs.append((s.primal(), s.info(), s.trace(10)))
        s.close()
    a, b = outs
    print(grid, "info", a[1], b[1], "x equal", np.array_equal(a[0], b[0]), np.abs(a[0] - b[0]).max())
    for k in a[2]:
        print("   ", k, a[2][k][0], b[2][k][0])

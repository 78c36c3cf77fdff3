import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polympc_b200
from polympc_b200 import workloads as W
from oracle import pyoracle
pmb = polympc_b200.load(); orc = pyoracle.load()
w = W.mobile_robot(32, sqp_max_iter=3, ls_max_iter=10)
outs = []
for api in (pmb, orc):
    s = api.sqp(w.name, 32); W.configure(s, w); s.solve()
    s.set_initial_conditions(w.x0 + 0.01); s.solve()
    outs.append((s.primal(), s.dual(), s.info(), s.trace(3), s.stats()))
    s.close()
a, b = outs
bad = np.argwhere(~((a[0] == b[0]) | (np.isnan(a[0]) & np.isnan(b[0]))))
print("bad x entries", bad, [(a[0][i, j], b[0][i, j]) for i, j in bad])
bad = np.argwhere(~((a[1] == b[1]) | (np.isnan(a[1]) & np.isnan(b[1]))))
print("bad lam entries", bad[:10], [(a[1][i, j], b[1][i, j]) for i, j in bad[:10]])
print("info", a[2][22], b[2][22]); print({k: (a[3][k][22], b[3][k][22]) for k in a[3]}); print(a[4][22], b[4][22])

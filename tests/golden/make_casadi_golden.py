"""Generate tests/golden/casadi_robot_5x2.npz from the REFERENCE's own CasADi-generated C fixtures.

The fixtures (reference tests/solvers/sqp/casadi_codegen/*.cpp, used by tests/solvers/sqp/codegen_test.cpp:61-437) are an
independent implementation of cost / constraints / gradients / Jacobian / Hessians of the mobile-robot transcription
(NX=3, NU=2, Chebyshev order 5 x 2 segments, t in [0,1], d=1, Q=R=I): 55 variables, 33 equality constraints.  They are
compiled from where they lie under /root/reference by oracle/Makefile into oracle/_ref/libcasadi_robot.so (never copied).
This script evaluates them on seeded inputs and stores dense outputs, so that the CPU test-suite can pin the oracle
without the reference tree being present (the GPU box has no /root/reference).

Run:  python tests/golden/make_casadi_golden.py      (in the build container, after `make -C oracle`)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "..", "oracle", "_ref", "libcasadi_robot.so")
c_ll = C.c_longlong


def casadi_call(lib, name, args, dense=True):
    fn = getattr(lib, name)
    work = getattr(lib, name + "_work")
    sp_out = getattr(lib, name + "_sparsity_out")
    sp_out.restype = C.POINTER(c_ll)
    sp_out.argtypes = [c_ll]
    sz = [c_ll(0) for _ in range(4)]
    work(*[C.byref(s) for s in sz])
    sz_arg, sz_res, sz_iw, sz_w = [int(s.value) for s in sz]
    sp = sp_out(0)
    nrow, ncol = int(sp[0]), int(sp[1])
    colind = [int(sp[2 + i]) for i in range(ncol + 1)]
    nnz = colind[-1]
    rows = [int(sp[3 + ncol + i]) for i in range(nnz)]
    argv = (C.POINTER(C.c_double) * max(sz_arg, len(args)))()
    keep = []
    for i, a in enumerate(args):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        argv[i] = a.ctypes.data_as(C.POINTER(C.c_double))
    out = np.zeros(nnz)
    resv = (C.POINTER(C.c_double) * max(sz_res, 1))()
    resv[0] = out.ctypes.data_as(C.POINTER(C.c_double))
    iw = (c_ll * max(sz_iw, 1))()
    w = (C.c_double * max(sz_w, 1))()
    rc = fn(argv, resv, iw, w, None)
    assert rc == 0
    full = np.zeros((nrow, ncol))
    for c in range(ncol):
        for k in range(colind[c], colind[c + 1]):
            full[rows[k], c] = out[k]
    return full


def main():
    lib = C.CDLL(LIB)
    rng = np.random.default_rng(20260117)
    cases = [np.ones(55), np.zeros(55)]
    lams = [np.ones(33), np.zeros(33)]
    for _ in range(6):
        x = rng.uniform(-1.0, 1.0, 55)
        x[33:] = rng.uniform(-1.5, 1.5, 22)
        cases.append(x)
        lams.append(rng.uniform(-2.0, 2.0, 33))
    X = np.stack(cases)
    L = np.stack(lams)
    out = dict(x=X, lam=L)
    res = {k: [] for k in ("cost", "cost_gradient", "cost_hessian", "constraints", "constraints_jacobian", "lagrangian",
                           "lagrangian_gradient", "lagrangian_hessian")}
    for x, lam in zip(X, L):
        res["cost"].append(casadi_call(lib, "fcost", [x]).ravel())
        res["cost_gradient"].append(casadi_call(lib, "fcost_gradient", [x]).ravel())
        res["cost_hessian"].append(casadi_call(lib, "fcost_hessian", [x]))
        res["constraints"].append(casadi_call(lib, "fconstraint", [x]).ravel())
        res["constraints_jacobian"].append(casadi_call(lib, "fconstraints_jacobian", [x]))
        res["lagrangian"].append(casadi_call(lib, "flagrangian", [x, lam]).ravel())
        res["lagrangian_gradient"].append(casadi_call(lib, "flagrangian_gradient", [x, lam]).ravel())
        res["lagrangian_hessian"].append(casadi_call(lib, "flagrangian_hessian", [x, lam]))
    for k, v in res.items():
        out[k] = np.stack(v)
    path = os.path.join(HERE, "casadi_robot_5x2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

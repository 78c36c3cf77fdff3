"""ctypes binding of the C ABI declared in include/polympc_b200.h.

This is harness plumbing for tests and bench.py (numpy in / numpy out); the product is the shared library.  The class
is parameterised by symbol prefix so that the CPU oracle (oracle/, prefix ``orc_``, test infrastructure) can be driven
through the very same wrapper — the parity tests then read identically on both sides.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("NX", "NU", "NP", "ND", "NG", "P", "S", "NN", "N", "M", "DUAL", "NPARAM")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SqpSettings(C.Structure):
    _fields_ = [("tau", C.c_double), ("eta", C.c_double), ("rho", C.c_double), ("eps_prim", C.c_double),
                ("eps_dual", C.c_double), ("max_iter", C.c_int), ("line_search_max_iter", C.c_int)]


class SqpInfo(C.Structure):
    _fields_ = [("iter", C.c_int), ("qp_solver_iter", C.c_int), ("status", C.c_int)]


class QpSettings(C.Structure):
    _fields_ = [("eps_rel", C.c_double), ("eps_abs", C.c_double), ("max_iter", C.c_int), ("warm_start", C.c_int),
                ("reuse_pattern", C.c_int), ("verbose", C.c_int), ("rho", C.c_double), ("sigma", C.c_double),
                ("alpha", C.c_double), ("check_termination", C.c_int), ("adaptive_rho", C.c_int),
                ("adaptive_rho_tolerance", C.c_double), ("adaptive_rho_interval", C.c_int), ("_pad", C.c_int)]


class QpInfo(C.Structure):
    _fields_ = [("status", C.c_int), ("iter", C.c_int), ("rho_updates", C.c_int), ("_pad", C.c_int),
                ("rho_estimate", C.c_double), ("res_prim", C.c_double), ("res_dual", C.c_double)]


QP_INFO_DTYPE = np.dtype([("status", "i4"), ("iter", "i4"), ("rho_updates", "i4"), ("_pad", "i4"),
                          ("rho_estimate", "f8"), ("res_prim", "f8"), ("res_dual", "f8")])
SQP_INFO_DTYPE = np.dtype([("iter", "i4"), ("qp_solver_iter", "i4"), ("status", "i4")])

SQP_SOLVED, SQP_MAX_ITER_EXCEEDED = 0, 1
QP_SOLVED, QP_MAX_ITER_EXCEEDED, QP_UNSOLVED = 0, 1, 2
INEQUALITY_CONSTRAINT, EQUALITY_CONSTRAINT, LOOSE_BOUNDS = 0, 1, 2

# every function name declared in include/polympc_b200.h (checked against the header by tests/test_abi.py)
ABI_FUNCTIONS = [
    "version", "last_error", "device_count", "problem_count", "problem_name", "problem_dims", "register_problem",
    "qp_default_settings", "sqp_default_settings", "sqp_default_qp_settings", "dm_eval", "cheb_tables",
    "ocp_create", "ocp_destroy", "ocp_dims", "ocp_set_params", "ocp_get_params", "ocp_set_time_limits", "ocp_time_nodes",
    "ocp_cost", "ocp_equalities", "ocp_inequalities", "ocp_equalities_linearised", "ocp_cost_gradient",
    "ocp_cost_gradient_hessian", "ocp_lagrangian_gradient", "ocp_lagrangian_gradient_hessian", "ocp_block_bfgs_update",
    "qp_solve", "qp_solve_admm", "kkt_assemble", "kkt_assemble_dev", "bfgs_update",
    "sqp_create", "sqp_destroy", "sqp_problem", "sqp_batch", "sqp_set_settings", "sqp_get_settings",
    "sqp_set_qp_settings", "sqp_get_qp_settings", "sqp_set_hessian_options", "sqp_set_hessian_update", "sqp_set_qp_solver", "sqp_set_preconditioner", "sqp_set_line_search", "sqp_set_filter", "sqp_get_filter", "ruiz_equilibrate", "ruiz_unscale", "sqp_set_trace", "sqp_set_schedule", "set_default_arithmetic", "get_default_arithmetic", "sqp_set_arithmetic", "sqp_get_arithmetic", "sqp_set_bounds_x", "sqp_set_bounds_g", "sqp_set_parameters",
    "sqp_set_primal", "sqp_set_dual", "sqp_set_initial_conditions", "sqp_reset_guess", "sqp_solve", "sqp_solve_async", "sqp_wait", "sqp_get_primal", "sqp_get_dual",
    "sqp_get_info", "sqp_get_stats", "sqp_get_trace", "sqp_last_solve_ms", "sqp_last_kernel_ms", "sqp_last_solve_launches", "sqp_set_profiling",
    "sqp_get_kernel_times", "sqp_get_phase_cycles", "sqp_set_stream",
]


FILTER_CAP = 16
FILTER_DOUBLES = 1 + 2 * FILTER_CAP
PRECOND_IDENTITY, PRECOND_RUIZ_DENSE, PRECOND_RUIZ_SPARSE = 0, 1, 2
LS_L1_MERIT, LS_FILTER = 0, 1


class PmbError(RuntimeError):
    pass


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _pi(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


class CApi:
    """Thin numpy wrapper over one shared library exporting ``<prefix>*``."""

    def __init__(self, lib: C.CDLL, prefix: str = "pmb_"):
        self.lib = lib
        self.prefix = prefix
        g = self._fn
        g("version").restype = C.c_char_p
        g("last_error").restype = C.c_char_p
        g("problem_name").restype = C.c_char_p
        g("problem_name").argtypes = [C.c_int]
        g("problem_dims").argtypes = [C.c_char_p, C.POINTER(Dims)]
        g("ocp_create").restype = C.c_void_p
        g("ocp_create").argtypes = [C.c_char_p, C.c_int]
        g("ocp_destroy").argtypes = [C.c_void_p]
        g("ocp_destroy").restype = None
        g("ocp_dims").argtypes = [C.c_void_p, C.POINTER(Dims)]
        g("ocp_set_params").argtypes = [C.c_void_p, c_double_p, C.c_int]
        g("ocp_get_params").argtypes = [C.c_void_p, c_double_p, C.c_int]
        g("ocp_set_time_limits").argtypes = [C.c_void_p, C.c_double, C.c_double]
        g("ocp_time_nodes").argtypes = [C.c_void_p, c_double_p]
        for name, nout in (("ocp_cost", 1), ("ocp_equalities", 1), ("ocp_inequalities", 1), ("ocp_equalities_linearised", 2),
                           ("ocp_cost_gradient", 2), ("ocp_cost_gradient_hessian", 3)):
            g(name).argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p] + [c_double_p] * nout
        g("ocp_lagrangian_gradient").argtypes = [C.c_void_p, C.c_int] + [c_double_p] * 8
        g("ocp_lagrangian_gradient_hessian").argtypes = [C.c_void_p, C.c_int] + [c_double_p] * 9
        g("cheb_tables").argtypes = [C.c_int, c_double_p, c_double_p, c_double_p]
        g("dm_eval").argtypes = [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]
        g("qp_solve").argtypes = [C.c_int, C.c_int, C.c_int] + [c_double_p] * 9 + [C.POINTER(QpSettings)] + \
            [c_double_p, c_double_p, C.c_void_p, c_double_p, c_double_p, c_int_p, c_int_p, c_int_p]
        g("qp_solve_admm").argtypes = [C.c_int, C.c_int, C.c_int] + [c_double_p] * 9 + [C.POINTER(QpSettings)] + \
            [c_double_p, c_double_p, C.c_void_p, c_double_p, c_int_p, c_int_p, c_int_p]
        g("kkt_assemble").argtypes = [C.c_int, C.c_int, C.c_int] + [c_double_p] * 4 + [C.c_double, c_double_p]
        g("bfgs_update").argtypes = [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_int_p]
        g("ocp_block_bfgs_update").argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, c_int_p]
        g("sqp_create").restype = C.c_void_p
        g("sqp_create").argtypes = [C.c_char_p, C.c_int, C.c_int]
        g("sqp_destroy").argtypes = [C.c_void_p]
        g("sqp_destroy").restype = None
        g("sqp_problem").restype = C.c_void_p
        g("sqp_problem").argtypes = [C.c_void_p]
        g("sqp_batch").argtypes = [C.c_void_p]
        g("sqp_set_settings").argtypes = [C.c_void_p, C.POINTER(SqpSettings)]
        g("sqp_get_settings").argtypes = [C.c_void_p, C.POINTER(SqpSettings)]
        g("sqp_set_qp_settings").argtypes = [C.c_void_p, C.POINTER(QpSettings)]
        g("sqp_get_qp_settings").argtypes = [C.c_void_p, C.POINTER(QpSettings)]
        g("sqp_set_hessian_options").argtypes = [C.c_void_p, C.c_int, C.c_int]
        g("sqp_set_hessian_update").argtypes = [C.c_void_p, C.c_int]
        g("sqp_set_preconditioner").argtypes = [C.c_void_p, C.c_int]
        g("sqp_set_qp_solver").argtypes = [C.c_void_p, C.c_int]
        g("sqp_set_line_search").argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        g("sqp_set_filter").argtypes = [C.c_void_p, c_double_p, C.c_int]
        g("sqp_get_filter").argtypes = [C.c_void_p, c_double_p]
        g("ruiz_equilibrate").argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [c_double_p] * 10
        g("ruiz_unscale").argtypes = [C.c_int, C.c_int, C.c_int] + [c_double_p] * 12
        g("sqp_set_trace").argtypes = [C.c_void_p, C.c_int]
        g("sqp_set_schedule").argtypes = [C.c_void_p, C.c_int]
        g("set_default_arithmetic").argtypes = [C.c_int]
        g("sqp_set_arithmetic").argtypes = [C.c_void_p, C.c_int]
        g("sqp_get_arithmetic").argtypes = [C.c_void_p]
        for name in ("sqp_set_bounds_x", "sqp_set_bounds_g"):
            g(name).argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_int]
        for name in ("sqp_set_parameters", "sqp_set_primal", "sqp_set_dual"):
            g(name).argtypes = [C.c_void_p, c_double_p, C.c_int]
        g("sqp_set_initial_conditions").argtypes = [C.c_void_p, c_double_p, c_double_p]
        g("sqp_solve").argtypes = [C.c_void_p]
        g("sqp_solve_async").argtypes = [C.c_void_p]
        g("sqp_wait").argtypes = [C.c_void_p]
        g("sqp_get_primal").argtypes = [C.c_void_p, c_double_p]
        g("sqp_get_dual").argtypes = [C.c_void_p, c_double_p]
        g("sqp_get_info").argtypes = [C.c_void_p, C.c_void_p]
        g("sqp_get_stats").argtypes = [C.c_void_p, c_double_p]
        g("sqp_get_trace").argtypes = [C.c_void_p, C.c_int, c_int_p, c_double_p, c_int_p, c_int_p, c_int_p]
        g("sqp_last_solve_ms").restype = C.c_double
        g("sqp_last_solve_ms").argtypes = [C.c_void_p]
        g("sqp_last_kernel_ms").restype = C.c_double
        g("sqp_last_kernel_ms").argtypes = [C.c_void_p]
        g("sqp_last_solve_launches").restype = C.c_longlong
        g("sqp_last_solve_launches").argtypes = [C.c_void_p]
        g("sqp_set_stream").argtypes = [C.c_void_p, C.c_void_p]
        g("sqp_reset_guess").argtypes = [C.c_void_p]
        g("sqp_set_profiling").argtypes = [C.c_void_p, C.c_int]
        g("sqp_get_kernel_times").argtypes = [C.c_void_p, c_double_p, C.POINTER(C.c_longlong)]
        g("sqp_get_phase_cycles").argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        g("kkt_assemble_dev").argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_double, C.c_void_p, C.c_void_p]
        g("qp_default_settings").argtypes = [C.POINTER(QpSettings)]
        g("qp_default_settings").restype = None
        g("sqp_default_settings").argtypes = [C.POINTER(SqpSettings)]
        g("sqp_default_settings").restype = None
        g("sqp_default_qp_settings").argtypes = [C.POINTER(QpSettings)]
        g("sqp_default_qp_settings").restype = None

    # -- plumbing
    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _chk(self, rc, what):
        if rc != 0:
            raise PmbError(f"{self.prefix}{what} failed with code {rc}: {self.last_error()}")

    def version(self):
        return self._fn("version")().decode()

    def last_error(self):
        s = self._fn("last_error")()
        return s.decode() if s else ""

    def device_count(self):
        return int(self._fn("device_count")())

    def problems(self):
        return [self._fn("problem_name")(i).decode() for i in range(self._fn("problem_count")())]

    def dims(self, name: str) -> dict:
        d = Dims()
        self._chk(self._fn("problem_dims")(name.encode(), C.byref(d)), "problem_dims")
        return d.as_dict()

    def qp_default_settings(self) -> QpSettings:
        s = QpSettings()
        self._fn("qp_default_settings")(C.byref(s))
        return s

    def sqp_default_settings(self) -> SqpSettings:
        s = SqpSettings()
        self._fn("sqp_default_settings")(C.byref(s))
        return s

    def sqp_default_qp_settings(self) -> QpSettings:
        s = QpSettings()
        self._fn("sqp_default_qp_settings")(C.byref(s))
        return s

    DM_FUNCTIONS = ("sin", "cos", "tan", "exp", "log", "atan2", "asin", "acos", "sinh", "cosh", "tanh", "pow", "sqrt")

    def dm_eval(self, fn: str, x, y=None):
        """deterministic elementary function `fn` of csrc/pmb_detmath.h applied element-wise"""
        x = _f64(x).ravel()
        yy = None if y is None else _f64(y).ravel()
        out = np.zeros_like(x)
        self._chk(self._fn("dm_eval")(self.DM_FUNCTIONS.index(fn), x.size, _p(x), _p(yy), _p(out)), "dm_eval")
        return out

    def cheb_tables(self, P: int):
        nodes = np.zeros(P + 1)
        D = np.zeros((P + 1) * (P + 1))
        w = np.zeros(P + 1)
        self._chk(self._fn("cheb_tables")(P, _p(nodes), _p(D), _p(w)), "cheb_tables")
        return nodes, D.reshape(P + 1, P + 1).T.copy(), w  # D returned as a [row, col] numpy array

    def ocp(self, name: str, device: int = 0) -> "Ocp":
        return Ocp(self, name, device)

    def sqp(self, name: str, batch: int, device: int = 0) -> "Sqp":
        return Sqp(self, name, batch, device)

    def set_default_arithmetic(self, mode: int):
        """process-wide default arithmetic: used by qp_solve and inherited by new Sqp handles (0 exact, 1 fast)"""
        self._chk(self._fn("set_default_arithmetic")(int(mode)), "set_default_arithmetic")

    # -- QP
    def qp_solve(self, H, h, A, Alb, Aub, xlb, xub, settings: QpSettings, x_guess=None, y_guess=None, extras=True):
        """H[b,N,N] (numpy [row,col]), h[b,N], A[b,M,N], bounds; returns dict of numpy arrays."""
        H = np.asarray(H, dtype=np.float64)
        B, N = H.shape[0], H.shape[1]
        A = np.asarray(A, dtype=np.float64)
        M = A.shape[1] if A.ndim == 3 else (A.size // max(1, B * N))
        A = A.reshape(B, M, N)
        Hc = np.ascontiguousarray(np.transpose(H, (0, 2, 1)))  # column-major per instance
        Ac = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
        h = _f64(h, (B, N)); Alb = _f64(Alb, (B, M)); Aub = _f64(Aub, (B, M)); xlb = _f64(xlb, (B, N)); xub = _f64(xub, (B, N))
        xg = None if x_guess is None else _f64(x_guess, (B, N))
        yg = None if y_guess is None else _f64(y_guess, (B, N + M))
        x = np.zeros((B, N)); y = np.zeros((B, N + M)); info = np.zeros(B, dtype=QP_INFO_DTYPE)
        z = np.zeros((B, M)) if extras else None
        q = np.zeros((B, N)) if extras else None
        perm = np.zeros((B, N + M), dtype=np.int32) if extras else None
        ctype = np.zeros((B, N + M), dtype=np.int32) if extras else None
        nf = np.zeros(B, dtype=np.int32) if extras else None
        rc = self._fn("qp_solve")(N, M, B, _p(Hc), _p(h), _p(Ac), _p(Alb), _p(Aub), _p(xlb), _p(xub), _p(xg), _p(yg),
                                  C.byref(settings), _p(x), _p(y), info.ctypes.data_as(C.c_void_p), _p(z), _p(q), _pi(perm),
                                  _pi(ctype), _pi(nf))
        self._chk(rc, "qp_solve")
        return dict(x=x, y=y, info=info, z=z, q=q, perm=perm, ctype=ctype, n_factor=nf)

    def qp_solve_admm(self, H, h, A, Alb, Aub, xlb, xub, settings: QpSettings, x_guess=None, y_guess=None, extras=True):
        """the OSQP-style ADMM<> of the reference (box rows appended to A); same conventions as qp_solve"""
        H = np.asarray(H, dtype=np.float64)
        B, N = H.shape[0], H.shape[1]
        A = np.asarray(A, dtype=np.float64)
        M = A.shape[1] if A.ndim == 3 else (A.size // max(1, B * N))
        A = A.reshape(B, M, N)
        Hc = np.ascontiguousarray(np.transpose(H, (0, 2, 1)))
        Ac = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
        h = _f64(h, (B, N)); Alb = _f64(Alb, (B, M)); Aub = _f64(Aub, (B, M)); xlb = _f64(xlb, (B, N)); xub = _f64(xub, (B, N))
        xg = None if x_guess is None else _f64(x_guess, (B, N))
        yg = None if y_guess is None else _f64(y_guess, (B, N + M))
        x = np.zeros((B, N)); y = np.zeros((B, N + M)); info = np.zeros(B, dtype=QP_INFO_DTYPE)
        z = np.zeros((B, M + N)) if extras else None
        perm = np.zeros((B, 2 * N + M), dtype=np.int32) if extras else None
        ctype = np.zeros((B, N + M), dtype=np.int32) if extras else None
        nf = np.zeros(B, dtype=np.int32) if extras else None
        rc = self._fn("qp_solve_admm")(N, M, B, _p(Hc), _p(h), _p(Ac), _p(Alb), _p(Aub), _p(xlb), _p(xub), _p(xg), _p(yg),
                                       C.byref(settings), _p(x), _p(y), info.ctypes.data_as(C.c_void_p), _p(z), _pi(perm), _pi(ctype), _pi(nf))
        self._chk(rc, "qp_solve_admm")
        return dict(x=x, y=y, info=info, z=z, perm=perm, ctype=ctype, n_factor=nf)

    def kkt_assemble(self, H, A, rho_box, rho_inv, sigma):
        H = np.asarray(H, dtype=np.float64)
        B, N = H.shape[0], H.shape[1]
        A = np.asarray(A, dtype=np.float64).reshape(B, -1, N)
        M = A.shape[1]
        Hc = np.ascontiguousarray(np.transpose(H, (0, 2, 1)))
        Ac = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
        K = np.zeros((B, N + M, N + M))
        self._chk(self._fn("kkt_assemble")(N, M, B, _p(Hc), _p(Ac), _p(_f64(rho_box, (B, N))), _p(_f64(rho_inv, (B, M))),
                                           float(sigma), _p(K)), "kkt_assemble")
        return np.transpose(K, (0, 2, 1)).copy()

    def ruiz_equilibrate(self, variant, H, h, A, Al, Au, l, u):
        """RuizEquilibration::compute on a batch (row-major numpy matrices in, scaled copies + D, E, c out)"""
        H = np.asarray(H, dtype=np.float64); A = np.asarray(A, dtype=np.float64)
        B, N, M = H.shape[0], H.shape[1], A.shape[1]
        Hc = np.ascontiguousarray(np.transpose(H, (0, 2, 1))); Ac = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
        v = [np.array(_f64(a, (B, n)), copy=True) for a, n in ((h, N), (Al, M), (Au, M), (l, N), (u, N))]
        D = np.zeros((B, N)); E = np.zeros((B, M)); c = np.zeros(B)
        self._chk(self._fn("ruiz_equilibrate")(N, M, B, int(variant), _p(Hc), _p(v[0]), _p(Ac), _p(v[1]), _p(v[2]), _p(v[3]), _p(v[4]),
                                               _p(D), _p(E), _p(c)), "ruiz_equilibrate")
        return dict(H=np.transpose(Hc, (0, 2, 1)).copy(), h=v[0], A=np.transpose(Ac, (0, 2, 1)).copy(), Al=v[1], Au=v[2], l=v[3], u=v[4], D=D, E=E, c=c)

    def ruiz_unscale(self, D, E, c, H, h, A, Al, Au, l, u, x=None, y=None):
        H = np.asarray(H, dtype=np.float64); A = np.asarray(A, dtype=np.float64)
        B, N, M = H.shape[0], H.shape[1], A.shape[1]
        Hc = np.ascontiguousarray(np.transpose(H, (0, 2, 1))); Ac = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
        v = [np.array(_f64(a, (B, n)), copy=True) for a, n in ((h, N), (Al, M), (Au, M), (l, N), (u, N))]
        xs = np.array(_f64(x, (B, N)), copy=True) if x is not None else None
        ys = np.array(_f64(y, (B, M + N)), copy=True) if y is not None else None
        self._chk(self._fn("ruiz_unscale")(N, M, B, _p(_f64(D, (B, N))), _p(_f64(E, (B, M))), _p(_f64(c, (B,))), _p(Hc), _p(v[0]), _p(Ac),
                                           _p(v[1]), _p(v[2]), _p(v[3]), _p(v[4]), _p(xs) if xs is not None else None,
                                           _p(ys) if ys is not None else None), "ruiz_unscale")
        return dict(H=np.transpose(Hc, (0, 2, 1)).copy(), h=v[0], A=np.transpose(Ac, (0, 2, 1)).copy(), Al=v[1], Au=v[2], l=v[3], u=v[4], x=xs, y=ys)

    def bfgs_update(self, Bm, s, y):
        Bm = np.asarray(Bm, dtype=np.float64)
        B, N = Bm.shape[0], Bm.shape[1]
        Bc = np.ascontiguousarray(np.transpose(Bm, (0, 2, 1)))
        branch = np.zeros(B, dtype=np.int32)
        self._chk(self._fn("bfgs_update")(N, B, _p(Bc), _p(_f64(s, (B, N))), _p(_f64(y, (B, N))), _pi(branch)), "bfgs_update")
        return np.transpose(Bc, (0, 2, 1)).copy(), branch


class Ocp:
    """ContinuousOCP transcription evaluations (batched)."""

    def __init__(self, api: CApi, name: str, device: int = 0, handle=None):
        self.api = api
        self._own = handle is None
        self.h = handle if handle is not None else api._fn("ocp_create")(name.encode(), device)
        if not self.h:
            raise PmbError(f"{api.prefix}ocp_create({name}) failed: {api.last_error()}")
        d = Dims()
        api._chk(api._fn("ocp_dims")(self.h, C.byref(d)), "ocp_dims")
        self.d = d.as_dict()

    def close(self):
        if self._own and self.h:
            self.api._fn("ocp_destroy")(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, v):
        v = _f64(v)
        self.api._chk(self.api._fn("ocp_set_params")(self.h, _p(v), v.size), "ocp_set_params")

    def get_params(self):
        v = np.zeros(self.d["NPARAM"])
        self.api._chk(self.api._fn("ocp_get_params")(self.h, _p(v), v.size), "ocp_get_params")
        return v

    def set_time_limits(self, t0, tf):
        self.api._chk(self.api._fn("ocp_set_time_limits")(self.h, float(t0), float(tf)), "ocp_set_time_limits")

    def time_nodes(self):
        t = np.zeros(self.d["NN"])
        self.api._chk(self.api._fn("ocp_time_nodes")(self.h, _p(t)), "ocp_time_nodes")
        return t

    def _in(self, var, d):
        D = self.d
        var = _f64(var).reshape(-1, D["N"])
        B = var.shape[0]
        dd = None if D["ND"] == 0 else _f64(d).reshape(B, D["ND"])
        return var, dd, B

    def cost(self, var, d=None):
        var, dd, B = self._in(var, d)
        c = np.zeros(B)
        self.api._chk(self.api._fn("ocp_cost")(self.h, B, _p(var), _p(dd), _p(c)), "ocp_cost")
        return c

    def equalities(self, var, d=None):
        var, dd, B = self._in(var, d)
        c = np.zeros((B, self.d["NX"] * self.d["NN"]))
        self.api._chk(self.api._fn("ocp_equalities")(self.h, B, _p(var), _p(dd), _p(c)), "ocp_equalities")
        return c

    def inequalities(self, var, d=None):
        var, dd, B = self._in(var, d)
        g = np.zeros((B, self.d["NG"] * self.d["NN"]))
        self.api._chk(self.api._fn("ocp_inequalities")(self.h, B, _p(var), _p(dd), _p(g)), "ocp_inequalities")
        return g

    def equalities_linearised(self, var, d=None):
        var, dd, B = self._in(var, d)
        ne, N = self.d["NX"] * self.d["NN"], self.d["N"]
        c = np.zeros((B, ne)); J = np.zeros((B, N, ne))
        self.api._chk(self.api._fn("ocp_equalities_linearised")(self.h, B, _p(var), _p(dd), _p(c), _p(J)), "ocp_equalities_linearised")
        return c, np.transpose(J, (0, 2, 1)).copy()

    def cost_gradient(self, var, d=None):
        var, dd, B = self._in(var, d)
        c = np.zeros(B); g = np.zeros((B, self.d["N"]))
        self.api._chk(self.api._fn("ocp_cost_gradient")(self.h, B, _p(var), _p(dd), _p(c), _p(g)), "ocp_cost_gradient")
        return c, g

    def cost_gradient_hessian(self, var, d=None):
        var, dd, B = self._in(var, d)
        N = self.d["N"]
        c = np.zeros(B); g = np.zeros((B, N)); H = np.zeros((B, N, N))
        self.api._chk(self.api._fn("ocp_cost_gradient_hessian")(self.h, B, _p(var), _p(dd), _p(c), _p(g), _p(H)), "ocp_cost_gradient_hessian")
        return c, g, np.transpose(H, (0, 2, 1)).copy()

    def lagrangian_gradient(self, var, lam, d=None):
        var, dd, B = self._in(var, d)
        D = self.d
        lam = _f64(lam).reshape(B, D["DUAL"])
        c = np.zeros(B); lg = np.zeros((B, D["N"])); cg = np.zeros((B, D["N"])); g = np.zeros((B, D["M"])); J = np.zeros((B, D["N"], D["M"]))
        self.api._chk(self.api._fn("ocp_lagrangian_gradient")(self.h, B, _p(var), _p(dd), _p(lam), _p(c), _p(lg), _p(cg), _p(g), _p(J)),
                      "ocp_lagrangian_gradient")
        return dict(cost=c, lag_grad=lg, cost_grad=cg, g=g, jac=np.transpose(J, (0, 2, 1)).copy())

    def lagrangian_gradient_hessian(self, var, lam, d=None):
        var, dd, B = self._in(var, d)
        D = self.d
        lam = _f64(lam).reshape(B, D["DUAL"])
        c = np.zeros(B); lg = np.zeros((B, D["N"])); cg = np.zeros((B, D["N"])); g = np.zeros((B, D["M"]))
        J = np.zeros((B, D["N"], D["M"])); H = np.zeros((B, D["N"], D["N"]))
        self.api._chk(self.api._fn("ocp_lagrangian_gradient_hessian")(self.h, B, _p(var), _p(dd), _p(lam), _p(c), _p(lg), _p(H), _p(cg),
                                                                      _p(g), _p(J)), "ocp_lagrangian_gradient_hessian")
        return dict(cost=c, lag_grad=lg, cost_grad=cg, g=g, jac=np.transpose(J, (0, 2, 1)).copy(),
                    hess=np.transpose(H, (0, 2, 1)).copy())

    def block_bfgs_update(self, Hm, s, y):
        """ContinuousOCP<..., SPARSE>::hessian_update_impl; Hm (B, N, N) row-major view of the Hessian -> (updated H, branch)"""
        Hm = _f64(Hm); B = Hm.shape[0]; N = self.d["N"]
        Hc = np.ascontiguousarray(np.transpose(Hm.reshape(B, N, N), (0, 2, 1)))      # column-major per instance
        branch = np.zeros(B, dtype=np.int32)
        self.api._chk(self.api._fn("ocp_block_bfgs_update")(self.h, B, _p(Hc), _p(_f64(s).reshape(B, N)), _p(_f64(y).reshape(B, N)), _pi(branch)),
                      "ocp_block_bfgs_update")
        return np.transpose(Hc, (0, 2, 1)).copy(), branch


class Sqp:
    """Batched SQPBase::solve with the MPC-facade style setters."""

    def __init__(self, api: CApi, name: str, batch: int, device: int = 0):
        self.api = api
        self.h = api._fn("sqp_create")(name.encode(), batch, device)
        if not self.h:
            raise PmbError(f"{api.prefix}sqp_create({name}, {batch}) failed: {api.last_error()}")
        self.batch = batch
        self.problem = Ocp(api, name, device, handle=api._fn("sqp_problem")(self.h))
        self.d = self.problem.d

    def close(self):
        if self.h:
            self.api._fn("sqp_destroy")(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # settings
    def settings(self) -> SqpSettings:
        s = SqpSettings()
        self.api._chk(self.api._fn("sqp_get_settings")(self.h, C.byref(s)), "sqp_get_settings")
        return s

    def set_settings(self, s: SqpSettings):
        self.api._chk(self.api._fn("sqp_set_settings")(self.h, C.byref(s)), "sqp_set_settings")

    def qp_settings(self) -> QpSettings:
        s = QpSettings()
        self.api._chk(self.api._fn("sqp_get_qp_settings")(self.h, C.byref(s)), "sqp_get_qp_settings")
        return s

    def set_qp_settings(self, s: QpSettings):
        self.api._chk(self.api._fn("sqp_set_qp_settings")(self.h, C.byref(s)), "sqp_set_qp_settings")

    def set_hessian_options(self, exact_every_iteration: bool = False, gershgorin_regularisation: bool = False):
        """the fixed menu of SQPBase CRTP overrides (reference tests/control/minimal_time_test.cpp:90-135)"""
        self.api._chk(self.api._fn("sqp_set_hessian_options")(self.h, int(exact_every_iteration), int(gershgorin_regularisation)),
                      "sqp_set_hessian_options")

    def set_qp_solver(self, kind: int):
        """0 boxADMM<> (default), 1 the OSQP-style ADMM<> (pmb_qp_solver_t)"""
        self.api._chk(self.api._fn("sqp_set_qp_solver")(self.h, int(kind)), "sqp_set_qp_solver")

    def set_preconditioner(self, kind: int):
        """0 identity, 1 RuizEquilibration<DENSE>, 2 RuizEquilibration<SPARSE> (pmb_preconditioner_t)"""
        self.api._chk(self.api._fn("sqp_set_preconditioner")(self.h, int(kind)), "sqp_set_preconditioner")

    def set_line_search(self, kind: int, beta: float = 1e-5, max_depth: int = 10):
        """0 l1 merit, 1 filter line search (pmb_line_search_t); empties the filters"""
        self.api._chk(self.api._fn("sqp_set_line_search")(self.h, int(kind), float(beta), int(max_depth)), "sqp_set_line_search")

    def set_filter(self, state):
        state = _f64(state)
        stride = 0 if state.ndim == 1 else state.shape[1]
        self.api._chk(self.api._fn("sqp_set_filter")(self.h, _p(state), stride), "sqp_set_filter")

    def filter(self):
        out = np.zeros((self.batch, FILTER_DOUBLES))
        self.api._chk(self.api._fn("sqp_get_filter")(self.h, _p(out)), "sqp_get_filter")
        return out

    def set_hessian_update(self, mode: int):
        """0 = dense damped BFGS (SQPBase default), 1 = the OCP's block BFGS (ContinuousOCP<..., SPARSE>::hessian_update_impl)"""
        self.api._chk(self.api._fn("sqp_set_hessian_update")(self.h, int(mode)), "sqp_set_hessian_update")

    def set_arithmetic(self, mode: int):
        """0 = PMB_ARITH_EXACT (bit-identical to the oracle), 1 = PMB_ARITH_FAST (fp64 tensor-core LDL^T, rounding-level differences)"""
        self.api._chk(self.api._fn("sqp_set_arithmetic")(self.h, int(mode)), "sqp_set_arithmetic")

    def set_schedule(self, schedule: int):
        """0 = FIFO, 1 = longest-processing-time-first from the previous solve's iteration counts (default)"""
        self.api._chk(self.api._fn("sqp_set_schedule")(self.h, int(schedule)), "sqp_set_schedule")

    def set_trace(self, on: bool = True):
        """record the per-iteration decision traces read by trace() (off by default)"""
        self.api._chk(self.api._fn("sqp_set_trace")(self.h, int(on)), "sqp_set_trace")

    def _vec(self, v, length):
        v = _f64(v)
        if v.size == length:
            return v.reshape(length), 0
        return v.reshape(self.batch, length), length

    def set_bounds_x(self, lbx, ubx):
        lb, st = self._vec(lbx, self.d["N"]); ub, st2 = self._vec(ubx, self.d["N"])
        assert st == st2
        self.api._chk(self.api._fn("sqp_set_bounds_x")(self.h, _p(lb), _p(ub), st), "sqp_set_bounds_x")

    def set_bounds_g(self, lbg, ubg):
        n = self.d["NG"] * self.d["NN"]
        lb, st = self._vec(lbg, n); ub, _ = self._vec(ubg, n)
        self.api._chk(self.api._fn("sqp_set_bounds_g")(self.h, _p(lb), _p(ub), st), "sqp_set_bounds_g")

    def set_parameters(self, d):
        v, st = self._vec(d, self.d["ND"])
        self.api._chk(self.api._fn("sqp_set_parameters")(self.h, _p(v), st), "sqp_set_parameters")

    def set_primal(self, x):
        v, st = self._vec(x, self.d["N"])
        self.api._chk(self.api._fn("sqp_set_primal")(self.h, _p(v), st), "sqp_set_primal")

    def set_dual(self, lam):
        v, st = self._vec(lam, self.d["DUAL"])
        self.api._chk(self.api._fn("sqp_set_dual")(self.h, _p(v), st), "sqp_set_dual")

    def set_initial_conditions(self, x0_lb, x0_ub=None):
        lb = _f64(x0_lb).reshape(self.batch, self.d["NX"])
        ub = lb if x0_ub is None else _f64(x0_ub).reshape(self.batch, self.d["NX"])
        self.api._chk(self.api._fn("sqp_set_initial_conditions")(self.h, _p(lb), _p(ub)), "sqp_set_initial_conditions")

    def set_stream(self, stream_ptr):
        self.api._chk(self.api._fn("sqp_set_stream")(self.h, C.c_void_p(stream_ptr)), "sqp_set_stream")

    def solve(self):
        self.api._chk(self.api._fn("sqp_solve")(self.h), "sqp_solve")

    def solve_async(self):
        """enqueue the batch solve on the handle's stream and return; wait() (or any getter) synchronises"""
        self.api._chk(self.api._fn("sqp_solve_async")(self.h), "sqp_solve_async")

    def wait(self):
        self.api._chk(self.api._fn("sqp_wait")(self.h), "sqp_wait")

    def reset_guess(self):
        self.api._chk(self.api._fn("sqp_reset_guess")(self.h), "sqp_reset_guess")

    def set_profiling(self, on: bool):
        self.api._chk(self.api._fn("sqp_set_profiling")(self.h, int(bool(on))), "sqp_set_profiling")

    def phase_cycles(self):
        """raw profiling counters of the last solve (see pmb_sqp_get_phase_cycles)"""
        cyc = np.zeros(16, dtype=np.uint64)
        self.api._chk(self.api._fn("sqp_get_phase_cycles")(self.h, cyc.ctypes.data_as(C.POINTER(C.c_ulonglong))), "sqp_get_phase_cycles")
        names = ("linearise", "qp", "step", "sqp_iterations", "qp_pivot", "qp_gather", "qp_factor", "qp_solve", "qp_update", "qp_resid", "admm_trips",
                 "_reserved", "fast_factor_diag", "fast_factor_panel", "fast_factor_trailing", "fast_factor_invert")
        return {k: int(cyc[i]) for i, k in enumerate(names) if not k.startswith("_")}

    def kernel_times(self):
        """{kernel name: (milliseconds, launches)} of the last solve (profiling must be on)"""
        ms = np.zeros(3)
        n = np.zeros(3, dtype=np.int64)
        self.api._chk(self.api._fn("sqp_get_kernel_times")(self.h, _p(ms), n.ctypes.data_as(C.POINTER(C.c_longlong))), "sqp_get_kernel_times")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("sqp_linearise", "qp_box_admm", "sqp_linesearch_step"))}

    def primal(self, out=None):
        """out: optional preallocated (e.g. pinned) float64 array of shape (batch, N)"""
        x = np.zeros((self.batch, self.d["N"])) if out is None else out
        self.api._chk(self.api._fn("sqp_get_primal")(self.h, _p(x)), "sqp_get_primal")
        return x

    def dual(self):
        lam = np.zeros((self.batch, self.d["DUAL"]))
        self.api._chk(self.api._fn("sqp_get_dual")(self.h, _p(lam)), "sqp_get_dual")
        return lam

    def info(self, out=None):
        info = np.zeros(self.batch, dtype=SQP_INFO_DTYPE) if out is None else out
        self.api._chk(self.api._fn("sqp_get_info")(self.h, info.ctypes.data_as(C.c_void_p)), "sqp_get_info")
        return info

    def stats(self):
        st = np.zeros((self.batch, 4))
        self.api._chk(self.api._fn("sqp_get_stats")(self.h, _p(st)), "sqp_get_stats")
        return st

    def trace(self, rows):
        qi = np.zeros((self.batch, rows), dtype=np.int32); al = np.zeros((self.batch, rows))
        bf = np.zeros((self.batch, rows), dtype=np.int32); ls = np.zeros((self.batch, rows), dtype=np.int32)
        qf = np.zeros((self.batch, rows), dtype=np.int32)
        self.api._chk(self.api._fn("sqp_get_trace")(self.h, rows, _pi(qi), _p(al), _pi(bf), _pi(ls), _pi(qf)), "sqp_get_trace")
        return dict(qp_iter=qi, alpha=al, bfgs=bf, ls_trials=ls, qp_factor=qf)

    def last_solve_ms(self):
        return float(self.api._fn("sqp_last_solve_ms")(self.h))

    def last_kernel_ms(self):
        return float(self.api._fn("sqp_last_kernel_ms")(self.h))

    def last_solve_launches(self):
        return int(self.api._fn("sqp_last_solve_launches")(self.h))

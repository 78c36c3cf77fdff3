"""Error behaviour of the C ABI (argument checking happens before any device work, so the warp-emulator build exercises the
same code on a CPU).  The reference reports problems through status enums and never throws; the C layer returns negative
pmb_error_t codes and keeps a message in pmb_last_error()."""
import ctypes as C

import numpy as np
import pytest

from polympc_b200.capi import PmbError


def test_unknown_problem(emu):
    with pytest.raises(PmbError):
        emu.sqp("no_such_problem", 4)
    with pytest.raises(PmbError):
        emu.ocp("no_such_problem")
    assert emu._fn("sqp_create")(b"mobile_robot_6x2", 0, 0) is None      # non-positive batch


def test_null_and_size_checks(emu):
    s = emu.sqp("mobile_robot_6x2", 2)
    f = emu._fn
    assert f("sqp_set_bounds_x")(s.h, None, None, 65) == -2              # PMB_ERR_BAD_ARGUMENT
    v = np.zeros(2 * 65)
    p = v.ctypes.data_as(C.POINTER(C.c_double))
    assert f("sqp_set_bounds_x")(s.h, p, p, 64) == -2                    # stride must be 0 or N
    assert f("sqp_set_initial_conditions")(s.h, None, None) == -2
    assert f("sqp_get_primal")(s.h, None) == -2
    assert f("sqp_solve")(None) == -2
    assert "bad" in emu.last_error() or "null" in emu.last_error()
    o = emu.ocp("mobile_robot_6x2")
    with pytest.raises(PmbError):
        o.set_params(np.zeros(3))                                         # NPARAM is 8
    assert f("ocp_cost")(o.h, 1, p, None, p) == -2                       # static parameters required (ND = 1)
    assert f("cheb_tables")(1, p, p, p) == -2
    st = emu.sqp_default_qp_settings()
    assert f("qp_solve")(0, 0, 1, p, p, p, p, p, p, p, None, None, C.byref(st), p, p, None, None, None, None, None, None) == -2
    s.close()


def test_preconditioner_and_line_search_arguments(emu):
    s = emu.sqp("mobile_robot_6x2", 2)
    f = emu._fn
    assert f("sqp_set_preconditioner")(s.h, 3) == -2 and f("sqp_set_preconditioner")(s.h, -1) == -2
    assert f("sqp_set_qp_solver")(s.h, 2) == -2 and f("sqp_set_qp_solver")(None, 0) == -2
    assert f("sqp_set_line_search")(s.h, 2, 0.1, 10) == -2
    assert f("sqp_set_line_search")(s.h, 1, 0.1, 0) == -2 and f("sqp_set_line_search")(s.h, 1, 0.1, 17) == -2     # depth in 1..PMB_FILTER_CAP
    assert f("sqp_set_line_search")(s.h, 1, float("nan"), 10) == -2
    buf = np.zeros((2, 33)); p = buf.ctypes.data_as(C.POINTER(C.c_double))
    assert f("sqp_get_filter")(s.h, p) == -5                             # PMB_ERR_UNSUPPORTED: the filter line search is off
    s.set_line_search(1, 0.1, 4)
    assert (s.filter() == 0).all()
    buf[:, 0] = 17
    assert f("sqp_set_filter")(s.h, p, 33) == -2                         # size out of range
    buf[:, 0] = 2; buf[:, 1:3] = [5.0, 6.0]; buf[:, 17:19] = [1.0, 0.5]
    s.set_filter(buf)
    assert np.array_equal(s.filter(), buf)
    s.set_line_search(1, 0.1, 4)                                         # selecting the line search again empties the filters
    assert (s.filter()[:, 0] == 0).all()
    d = np.zeros(4); pd = d.ctypes.data_as(C.POINTER(C.c_double))
    assert f("ruiz_equilibrate")(2, 1, 1, 0, pd, pd, pd, pd, pd, pd, pd, pd, pd, pd) == -2      # variant must be a Ruiz variant
    assert f("ruiz_equilibrate")(2, 1, 1, 1, None, pd, pd, pd, pd, pd, pd, pd, pd, pd) == -2
    s.close()


def test_zero_max_iter_and_empty_work(emu, orc):
    """max_iter = 0: like the reference (sqp_base.hpp:583-637 precede the while loop) exactly one iteration is done"""
    from polympc_b200 import workloads as W
    w = W.mobile_robot(2, sqp_max_iter=0)
    out = []
    for api in (emu, orc):
        s = api.sqp(w.name, 2); W.configure(s, w); s.solve()
        info = s.info()
        assert (info["iter"] == 1).all() and (info["status"] == 1).all() and (info["qp_solver_iter"] > 0).all()
        out.append(s.primal())
        s.close()
    assert np.array_equal(out[0], out[1])


def test_required_outputs_and_borrowed_handles(emu):
    """a NULL output that the kernel of that operator writes unconditionally is refused with PMB_ERR_BAD_ARGUMENT (it used to be
    a device null write); the problem handle borrowed from an SQP handle survives pmb_ocp_destroy; the new switches check their
    arguments"""
    f = emu._fn
    o = emu.ocp("mobile_robot_6x2")
    v = np.zeros(65 * 104)
    p = v.ctypes.data_as(C.POINTER(C.c_double))
    assert f("ocp_cost_gradient")(o.h, 1, p, p, p, None) == -2
    assert f("ocp_equalities")(o.h, 1, p, p, None) == -2
    assert f("ocp_equalities_linearised")(o.h, 1, p, p, p, None) == -2
    assert f("ocp_lagrangian_gradient")(o.h, 1, p, p, p, p, None, p, p, p) == -2
    assert f("ocp_lagrangian_gradient_hessian")(o.h, 1, p, p, p, p, p, None, p, p, p) == -2
    assert f("ocp_block_bfgs_update")(o.h, 1, None, p, p, None) == -2
    s = emu.sqp("mobile_robot_6x2", 2)
    f("sqp_problem").restype = C.c_void_p
    f("ocp_destroy").argtypes = [C.c_void_p]
    borrowed = f("sqp_problem")(s.h)
    f("ocp_destroy")(borrowed)                                            # ignored: the handle belongs to s
    assert s.d["N"] == 65
    st = s.settings(); st.max_iter = 1; st.line_search_max_iter = 2; s.set_settings(st)
    s.solve()
    assert f("sqp_set_hessian_update")(s.h, 7) == -2
    assert f("sqp_set_arithmetic")(s.h, 7) == -2
    assert f("sqp_set_schedule")(s.h, 7) == -2
    assert f("set_default_arithmetic")(7) == -2
    with pytest.raises(PmbError):
        s.trace(4)                                                        # traces were off during the solve
    s.close()

"""csrc/pmb_detmath.h — the deterministic fp64 elementary functions every functor and the Chebyshev tables use.
CPU: accuracy of the host instantiation against numpy/libm.  GPU: the device instantiation must return the SAME BITS as the
host instantiation (this is what makes the data-dependent decisions of the SQP path reproducible across the two sides)."""
import numpy as np
import pytest

UNARY = {
    "sin": (np.sin, (-50.0, 50.0)), "cos": (np.cos, (-50.0, 50.0)), "tan": (np.tan, (-1.5, 1.5)), "exp": (np.exp, (-700.0, 700.0)),
    "log": (np.log, (1e-300, 1e300)), "asin": (np.arcsin, (-1.0, 1.0)), "acos": (np.arccos, (-1.0, 1.0)),
    "sinh": (np.sinh, (-20.0, 20.0)), "cosh": (np.cosh, (-20.0, 20.0)), "tanh": (np.tanh, (-20.0, 20.0)), "sqrt": (np.sqrt, (1e-300, 1e300)),
}


def ulp_err(a, ref):
    a = np.asarray(a, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return np.abs(a - ref) / np.maximum(np.spacing(np.abs(ref)), np.finfo(np.float64).tiny)


def samples(lo, hi, n, rng, log=False):
    if log:
        return np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    return rng.uniform(lo, hi, n)


@pytest.mark.parametrize("fn", sorted(UNARY))
def test_accuracy_vs_libm(orc, fn):
    ref, (lo, hi) = UNARY[fn]
    rng = np.random.default_rng(abs(hash(fn)) % 1000)
    x = samples(lo, hi, 20000, rng, log=fn in ("log", "sqrt"))
    got = orc.dm_eval(fn, x)
    tol = 0.5 if fn == "sqrt" else 2.0          # sqrt is correctly rounded; the others within 2 ulp of libm
    assert ulp_err(got, ref(x)).max() <= tol + 1.0, fn    # libm itself is within ~1 ulp of the true value


def test_accuracy_binary(orc):
    rng = np.random.default_rng(7)
    y, x = rng.uniform(-5, 5, 20000), rng.uniform(-5, 5, 20000)
    assert ulp_err(orc.dm_eval("atan2", y, x), np.arctan2(y, x)).max() <= 3.0
    b, e = rng.uniform(0.1, 30.0, 20000), rng.uniform(-8, 8, 20000)
    assert np.max(np.abs(orc.dm_eval("pow", b, e) / np.power(b, e) - 1.0)) <= 1e-13


def test_special_values(orc):
    inf, nan = np.inf, np.nan
    assert np.isnan(orc.dm_eval("sin", [inf, nan])).all() and orc.dm_eval("sin", [0.0])[0] == 0.0
    assert orc.dm_eval("cos", [0.0])[0] == 1.0
    e = orc.dm_eval("exp", [-inf, 0.0, 710.0, -750.0])
    assert e[0] == 0.0 and e[1] == 1.0 and e[2] == inf and e[3] == 0.0
    l = orc.dm_eval("log", [0.0, 1.0, -1.0, inf])
    assert l[0] == -inf and l[1] == 0.0 and np.isnan(l[2]) and l[3] == inf
    assert orc.dm_eval("sqrt", [4.0])[0] == 2.0 and np.isnan(orc.dm_eval("sqrt", [-1.0])[0])
    a = orc.dm_eval("atan2", [0.0, 1.0, -1.0, 0.0], [1.0, 0.0, 0.0, -1.0])
    assert a[0] == 0.0 and abs(a[1] - np.pi / 2) < 1e-15 and abs(a[2] + np.pi / 2) < 1e-15 and abs(a[3] - np.pi) < 1e-15
    # large arguments: accurate while the spacing of doubles is small against 2 pi, bounded beyond (exact folding modulo
    # fl(2^18 * 2 pi) before the Cody-Waite reduction)
    mid = np.array([1e5, 1e6, 3e6, 1e8, 1e9])
    assert np.abs(orc.dm_eval("sin", mid) - np.sin(mid)).max() < 1e-6
    assert np.abs(orc.dm_eval("cos", mid) - np.cos(mid)).max() < 1e-6
    big = np.array([1e15, 1e22, 1e100, 1e300, -1e250])
    assert np.abs(orc.dm_eval("sin", big)).max() <= 1.0 and np.abs(orc.dm_eval("cos", big)).max() <= 1.0
    s2 = orc.dm_eval("sin", big) ** 2 + orc.dm_eval("cos", big) ** 2
    assert np.abs(s2 - 1.0).max() < 1e-12


def test_emulator_build_matches_host(emu, orc):
    rng = np.random.default_rng(3)
    x = rng.uniform(-30, 30, 5000)
    for fn in ("sin", "cos", "exp", "tanh"):
        assert np.array_equal(emu.dm_eval(fn, x), orc.dm_eval(fn, x))


@pytest.mark.gpu
@pytest.mark.parametrize("fn", sorted(UNARY))
def test_device_bits_equal_host_bits(pmb, orc, fn):
    _, (lo, hi) = UNARY[fn]
    rng = np.random.default_rng(11)
    x = np.concatenate([samples(lo, hi, 400000, rng, log=fn in ("log", "sqrt")),
                        rng.uniform(-1e-300, 1e-300, 1000),                      # denormal neighbourhood
                        rng.standard_normal(100000) * 10.0 ** rng.integers(-20, 20, 100000),   # wide dynamic range, both signs
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 5e-324, 1.7976931348623157e308]])
    a, b = pmb.dm_eval(fn, x), orc.dm_eval(fn, x)
    same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{fn}: {np.count_nonzero(~same)} of {x.size} results differ between device and host"


@pytest.mark.gpu
def test_device_bits_equal_host_bits_binary(pmb, orc):
    rng = np.random.default_rng(12)
    y = rng.standard_normal(300000) * 10.0 ** rng.integers(-10, 10, 300000)
    x = rng.standard_normal(300000) * 10.0 ** rng.integers(-10, 10, 300000)
    for fn, (p, q) in (("atan2", (y, x)), ("pow", (np.abs(y) + 1e-3, np.clip(x, -50, 50)))):
        a, b = pmb.dm_eval(fn, p, q), orc.dm_eval(fn, p, q)
        same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), fn

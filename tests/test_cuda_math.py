"""Accuracy of the device math functions (csrc/soil_math.cuh) against numpy on the
argument ranges the closures produce.  Tolerance in ulp, written per function."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(kind, x, y=None):
    import climaland_b200 as cl
    L = cl._lib.lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    yp = np.ascontiguousarray(y, dtype=np.float64).ctypes.data_as(dp) if y is not None else None
    cl._lib.check(L.clb_test_math(kind, x.ctypes.data_as(dp), yp, out.ctypes.data_as(dp), x.size))
    return out


def _ulp_err(got, want):
    want = np.asarray(want, dtype=np.float64)
    return np.max(np.abs(got - want) / np.spacing(np.abs(want)))


def _samples(lo, hi, n=400000, log=True, seed=0):
    rng = np.random.default_rng(seed)
    if log:
        return np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    return rng.uniform(lo, hi, n)


def test_rcp_and_div():
    x = np.concatenate([_samples(1e-30, 1e30), -_samples(1e-12, 1e12), [1.0, 2.0, 0.5, 3.0, 1e-3]])
    assert _ulp_err(_run(0, x), 1.0 / x) <= 1.0
    a = np.concatenate([_samples(1e-20, 1e20, seed=1), [0.3, 0.124, 1.0, 7.0, 1e-3]])
    b = np.concatenate([_samples(1e-20, 1e20, seed=2), [0.3, 0.124, 1.0, 7.0, 1e-3]])
    q = _run(1, a, b)
    assert _ulp_err(q, a / b) <= 1.0
    assert np.all(q[-5:] == 1.0), "exact quotients must be exact (S == 1 branch parity)"
    seed = _run(5, x)
    print("rcp seed max rel err: 2^%.1f" % np.log2(np.max(np.abs(seed * x - 1.0))))


def test_div_by_is_the_ieee_quotient():
    """fm::div_by (division by a launch constant through its host-divided reciprocal): the bits of a / b"""
    a = np.concatenate([_samples(1e-20, 1e20, seed=11), -_samples(1e-9, 1e3, seed=12), [0.0, 0.3, 1.0, 7.0, np.inf]])
    b = np.concatenate([_samples(1e-3, 1e2, seed=13), _samples(1e-3, 1e2, seed=14), [0.025, 0.3, 3.0, 0.7, 0.025]])
    q = _run(8, a, b)
    assert np.array_equal(q, a / b)


def test_log():
    x = np.concatenate([_samples(1e-300, 1e300), _samples(1e-9, 1.0, seed=3), 1.0 - _samples(1e-16, 0.5, seed=4),
                        1.0 + _samples(1e-16, 0.5, seed=5), [1.0, 0.5, 2.0]])
    got = _run(2, x)
    assert _ulp_err(got[x != 1.0], np.log(x[x != 1.0])) <= 2.0
    assert got[-3] == 0.0
    sp = _run(2, np.array([0.0, -1.0, np.nan]))
    assert sp[0] == -np.inf and np.isnan(sp[1]) and np.isnan(sp[2])


def test_exp():
    x = np.concatenate([_samples(-700.0, 700.0, log=False), _samples(-1.0, 1.0, log=False, seed=6),
                        -_samples(1e-17, 1e-3, seed=7), [0.0]])
    got = _run(3, x)
    assert _ulp_err(got, np.exp(x)) <= 2.0
    sp = _run(3, np.array([-np.inf, -800.0, 800.0, np.inf, np.nan]))
    assert sp[0] == 0.0 and sp[1] == 0.0 and sp[2] == np.inf and sp[3] == np.inf and np.isnan(sp[4])


def test_sqrt():
    x = np.concatenate([_samples(1e-300, 1e300), _samples(1e-9, 1.0, seed=8), [0.0, 1.0, 4.0, 0.25]])
    got = _run(4, x)
    assert _ulp_err(got[x > 0], np.sqrt(x[x > 0])) <= 1.0
    assert got[-4] == 0.0 and got[-3] == 1.0 and got[-2] == 2.0 and got[-1] == 0.5


def test_table_log():
    """log_tab (soil_mathv.cuh): one table look-up + degree-6 series.  The closures only ever feed a
    logarithm into an exp, so the contract is absolute near x = 1: |err| <= 2 ulp or 2^-58."""
    x = np.concatenate([_samples(1e-300, 1e300), _samples(1e-9, 1.0, seed=3), 1.0 - _samples(1e-16, 0.5, seed=4),
                        1.0 + _samples(1e-16, 0.5, seed=5), [1.0, 0.5, 2.0, 0.6875, 1.375, np.nextafter(0.6875, 0)]])
    got = _run(6, x)
    want = np.log(x)
    err = np.abs(got - want)
    bound = np.maximum(2.0 * np.spacing(np.abs(want)), 2.0 ** -58)
    assert np.all(err <= bound), (x[np.argmax(err / bound)], np.max(err / bound))


def test_table_exp():
    x = np.concatenate([_samples(-700.0, 700.0, log=False), _samples(-1.0, 1.0, log=False, seed=6),
                        -_samples(1e-17, 1e-3, seed=7), [0.0]])
    got = _run(7, x)
    assert _ulp_err(got, np.exp(x)) <= 2.0
    assert got[-1] == 1.0

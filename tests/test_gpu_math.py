"""The pinned device math (det_exp / det_log / det_digamma) must equal the oracle's bit for bit,
including the range edges; everything downstream relies on it."""
import numpy as np
import pytest

import orc
import mmsig

pytestmark = pytest.mark.gpu


def _dev(fn, x):
    h = mmsig.capi.Handle()
    y = np.empty_like(x)
    h.check(h.lib.mmsig_debug_math(h.h, fn, x.size, mmsig.capi.dp(x), mmsig.capi.dp(y)))
    h.close()
    return y


def test_exp_log_digamma_bit_exact():
    L = orc.lib()
    rng = np.random.default_rng(0)
    xe = np.concatenate([rng.uniform(-30, 30, 200000), rng.uniform(-760, 720, 50000), rng.uniform(-1e-3, 1e-3, 10000),
                         [0.0, -0.0, 709.782712893384, 709.7827128933841, 709.78, 710.0, 1e3, np.inf, -745.13, -745.14,
                          -745.2, -745.3, -746.0, -1e3, -np.inf, -708.4, -744.0]])
    ye = _dev(0, xe)
    ref = np.array([L.orc_exp(float(v)) for v in xe])
    assert np.array_equal(ye, ref), xe[ye != ref][:10]
    assert np.isnan(_dev(0, np.array([np.nan]))[0])
    xl = np.concatenate([np.exp(rng.uniform(-700, 700, 100000)), rng.uniform(0.5, 2.0, 100000), rng.uniform(1e-7, 1e-2, 50000),
                         [1.0, 5e-324, 2.2250738585072014e-308, 1e-310, 1.7976931348623157e308, np.inf, 0.0, 2.0 ** 0.5]])
    yl = _dev(1, xl)
    ref = np.array([L.orc_log(float(v)) for v in xl])
    assert np.array_equal(yl, ref)
    assert np.isnan(_dev(1, np.array([-1.0]))[0]) and np.isnan(_dev(1, np.array([np.nan]))[0])
    xd = np.concatenate([rng.uniform(1e-3, 10, 100000), rng.uniform(1, 1e7, 100000), [0.1, 1.0, 6.999999, 7.0]])
    yd = _dev(2, xd)
    ref = np.array([L.orc_digamma_det(float(v)) for v in xd])
    assert np.array_equal(yd, ref)


def test_branch_free_div_rcp_sqrt_are_ieee():
    """fast_div / fast_rcp / fast_sqrt (det_math.cuh) are the compiler's correctly rounded
    sequences minus the exceptional-operand branch: on their stated domain they must equal IEEE
    division / square root (numpy on the host) bit for bit."""
    rng = np.random.default_rng(1)
    n = 1 << 20
    sign = lambda k: rng.choice([-1.0, 1.0], k)
    num = np.concatenate([sign(n) * np.exp(rng.uniform(-600, 600, n)), sign(n) * rng.uniform(0, 2, n), np.zeros(64), -np.zeros(64),
                          sign(n) * np.exp(rng.uniform(-40, 40, n))])
    den = np.concatenate([sign(n) * np.exp(rng.uniform(-60, 60, n)), -1.0 - rng.uniform(0, 1, n), sign(128) * rng.uniform(1e-6, 3, 128),
                          sign(n) * np.exp(rng.uniform(-40, 40, n))])
    # mantissa patterns near 1 and 2 (the hard cases of Newton division)
    eps = 2.0 ** -52
    hard = np.concatenate([1.0 + eps * np.arange(0, 4096), 2.0 - eps * np.arange(1, 4097), 1.5 + eps * np.arange(-2048, 2048)])
    num = np.concatenate([num, rng.permutation(hard), hard])
    den = np.concatenate([den, hard, rng.permutation(hard)])
    h = mmsig.capi.Handle()
    q = np.empty_like(num)
    x = np.ascontiguousarray(np.concatenate([num, den]))
    h.check(h.lib.mmsig_debug_math(h.h, 3, num.size, mmsig.capi.dp(x), mmsig.capi.dp(q)))
    h.close()
    ref = num / den
    # a zero numerator yields a zero whose sign may differ from IEEE's (-0 / d -> +0): harmless on
    # this path (the quotient is only ever added to a non-zero value or squared)
    bad = ~((q == ref) & ((np.signbit(q) == np.signbit(ref)) | (num == 0)))
    assert not bad.any(), (num[bad][:5], den[bad][:5], q[bad][:5], ref[bad][:5])
    d = np.concatenate([sign(n) * np.exp(rng.uniform(-600, 600, n)), rng.uniform(1e-7, 4, n), hard, -hard])
    r = _dev(4, d)
    assert np.array_equal(r, 1.0 / d), d[r != 1.0 / d][:5]
    xs = np.concatenate([np.exp(rng.uniform(-600, 600, n)), rng.uniform(0, 1, n), 2.0 ** -53 * np.arange(1, 4097), hard, hard * 2, hard * 0.5,
                         1.0 - 2.0 ** -53 * np.arange(1, 4097)])
    s = _dev(5, xs)
    assert np.array_equal(s, np.sqrt(xs)), xs[s != np.sqrt(xs)][:5]

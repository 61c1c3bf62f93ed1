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

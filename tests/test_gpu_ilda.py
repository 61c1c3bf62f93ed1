"""CUDA ILDA path (reference src/ILDA.jl, through the C ABI) against the oracle's ILDA, itself pinned
on the known answers of test/ilda.jl (tests/test_oracle_ilda.py).  Tolerances as for the LDA
(tests/test_gpu_lda.py): 1e-12 relative per iteration, 1e-11 on the ELBO."""
import json
import os

import numpy as np
import pytest

import orc
import mmsig
from mmsig.counts import from_nested
from conftest import ROOT
from util import small_synth, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12
G = json.load(open(os.path.join(ROOT, "tests", "golden", "ilda_known_answers.json")))


def _grid_features(shape):
    """every term = one combination of feature values (V = prod(shape))"""
    return np.stack(np.meshgrid(*[range(s) for s in shape], indexing="ij"), -1).reshape(-1, len(shape))


def _pair(K, alpha, eta, feat, csr, l0, arith=orc.ARITH_LITERAL):
    o = orc.OracleILDA(K, alpha, eta, feat, csr, l0, arith=arith, nthreads=8)
    g = mmsig.ILDA(K, alpha, eta, feat, csr, lambdaf0=l0)
    return o, g


def _check(o, g, ll_o=None, ll_g=None, tol=TOL, ctor=False):
    s = g.state()
    lf, ef = g.tables()
    assert rel_err(lf, o.lambdaf) <= tol
    assert np.max(np.abs(ef - o.Elnbetaf)) <= tol * 50
    assert rel_err(s["gamma"], o.gamma) <= tol
    assert rel_err(s["beta"], o.beta) <= tol
    assert np.max(np.abs(s["Elnbeta"] - o.Elnbeta)) <= tol * 100
    if not ctor:                                  # θ exists from the first update_θ! on (src/ILDA.jl:101-103)
        assert rel_err(s["theta"], o.theta) <= tol
        assert np.max(np.abs(s["Elntheta"] - o.Elntheta)) <= tol * 50
    if ll_o is not None:
        assert abs(ll_g - ll_o) <= tol * abs(ll_o)


def test_ilda_toy_of_the_reference_tests():
    feat = np.asarray(G["features"]) - 1
    csr = from_nested([[np.asarray(x)] for x in G["X"]], 1)[0]
    l0 = np.array([3., 50, 17, 99, 8, 21, 64, 5])
    o, g = _pair(G["K"], G["alpha"], G["eta"], feat, csr, l0)
    assert g.I == G["ctor"]["I"] and g.J == G["ctor"]["J"]
    _check(o, g, ctor=True)
    for _ in range(3):
        ll_o, ll_g = o.iterate(), g.iterate()
        _check(o, g, ll_o, ll_g)
    np.testing.assert_allclose(g.phi(), o.phi, rtol=1e-12)
    eo, to = o.elbo()
    eg, tg = g.calculate_elbo()
    np.testing.assert_allclose(tg, to, rtol=1e-10, atol=1e-10 * abs(eo))
    assert abs(eg - eo) <= 1e-10 * abs(eo)
    g.close()


@pytest.mark.parametrize("arith", [orc.ARITH_LITERAL, orc.ARITH_DET])
@pytest.mark.parametrize("K,shape,D,eta", [(5, (4, 4, 6), 2000, 0.1), (3, (3, 4), 300, [0.2, 0.7]), (32, (2, 3), 150, 0.1)])
def test_ilda_synthetic(K, shape, D, eta, arith):
    feat = _grid_features(shape)
    V = feat.shape[0]
    csr = small_synth(D, [K], [V], empty_frac=0.05)[0]
    l0 = np.random.default_rng(7).integers(1, 101, K * sum(shape)).astype(float)
    o, g = _pair(K, 0.1, eta, feat, csr, l0, arith=arith)
    for _ in range(4):
        ll_o, ll_g = o.iterate(), g.iterate()
        _check(o, g, ll_o, ll_g)
    eo, to = o.elbo()
    eg, tg = g.calculate_elbo()
    # ElnQβ is a difference of lgammas of O(|ELBO|): absolute tolerance on the ELBO's scale
    np.testing.assert_allclose(tg, to, rtol=1e-10, atol=1e-11 * abs(eo))
    assert abs(eg - eo) <= 1e-11 * abs(eo), (tg, to)
    g.close()


def test_ilda_fit_matches_oracle_fit():
    feat = _grid_features((4, 6))
    csr = small_synth(800, [4], [24])[0]
    l0 = np.random.default_rng(3).integers(1, 101, 4 * 10).astype(float)
    o, g = _pair(4, 0.1, 0.1, feat, csr, l0)
    ho = o.fit(maxiter=20)
    hg = g.fit(maxiter=20, verbose=False)
    assert len(hg) == len(ho)
    np.testing.assert_allclose(hg, ho, rtol=1e-10)
    assert abs(g.elbo - o.elbo()[0]) <= 1e-9 * abs(g.elbo)
    g.close()


def test_one_feature_ilda_is_the_lda():
    K, V, D = 4, 20, 600
    csr = small_synth(D, [K], [V])[0]
    l0 = np.random.default_rng(5).integers(1, 101, K * V).astype(float)
    a = mmsig.ILDA(K, 0.1, 0.1, np.arange(V).reshape(V, 1), csr, lambdaf0=l0)
    b = mmsig.LDA(K, 0.1, 0.1, csr, V=V, lambda0=l0)
    for _ in range(3):
        la, lb = a.iterate(), b.iterate()
        assert abs(la - lb) <= TOL * abs(lb)
    sa, sb = a.state(), b.state()
    assert rel_err(a.tables()[0], sb["lam"].ravel()) <= TOL
    assert rel_err(sa["beta"], sb["beta"]) <= TOL and rel_err(sa["gamma"], sb["gamma"]) <= TOL
    assert abs(a.calculate_elbo()[0] - b.calculate_elbo()[0]) <= 1e-11 * abs(b.calculate_elbo()[0])
    a.close()
    b.close()


def test_ilda_fit_heldout():
    feat = _grid_features((3, 5))
    K, V = 3, 15
    train = small_synth(500, [K], [V])[0]
    held = small_synth(120, [K], [V], seed=11)[0]
    l0 = np.random.default_rng(2).integers(1, 101, K * 8).astype(float)
    o, g = _pair(K, 0.1, 0.1, feat, train, l0)
    for _ in range(5):
        o.iterate(), g.iterate()
    oh = orc.OracleILDA(K, 0.1, 0.1, feat, held, o.lambdaf.copy(), nthreads=4)
    gh = g.fit_heldout(held, maxiter=12)
    ll_o = [oh.iterate_flags(orc.FLAG_FREEZE_TOPICS) for _ in range(len(gh.ll_history))]
    np.testing.assert_allclose(gh.ll_history, ll_o, rtol=1e-10)
    _check(oh, gh, tol=1e-10)
    assert np.array_equal(gh.tables()[0], g.tables()[0])            # feature tables stay frozen
    g.close()
    gh.close()


def test_ilda_argument_errors():
    feat = _grid_features((2, 3))
    csr = small_synth(50, [2], [6])[0]
    g = mmsig.ILDA(2, 0.1, 0.1, feat, csr, lambdaf0=np.ones(2 * 5))
    with pytest.raises(mmsig.capi.MmsigError):
        g.h.check(g.h.lib.mmsig_ilda_set_features(g.h.h, 2, np.ascontiguousarray(feat, np.int32).ctypes.data_as(mmsig.capi.c_i32p)))
    with pytest.raises(NotImplementedError):
        g.transform(csr)
    g.close()
    bad = feat.copy()
    bad[0, 0] = -1
    with pytest.raises(mmsig.capi.MmsigError):
        mmsig.ILDA(2, 0.1, 0.1, bad, csr, lambdaf0=np.ones(2 * 5))
    with pytest.raises(mmsig.capi.MmsigError):
        mmsig.ILDA(2, 0.1, [0.1, 0.0], feat, csr, lambdaf0=np.ones(2 * 5))
    plain = mmsig.LDA(2, 0.1, 0.1, csr, V=6, lambda0=np.ones(12))
    with pytest.raises(mmsig.capi.MmsigError):
        plain.h.check(plain.h.lib.mmsig_ilda_get_tables(plain.h.h, None, None))
    plain.close()

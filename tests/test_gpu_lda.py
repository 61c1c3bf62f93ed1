"""CUDA LDA path (through the C ABI) against the oracle.  Tolerance: 1e-12 relative per
iteration (north_star), 1e-8 relative ELBO for a fixed-iteration fit.  LDA has no data-dependent
branches, so the hoisted-exponential kernel is compared with plain tolerances, not bit-exactly."""
import numpy as np
import pytest

import orc
import mmsig
from mmsig.counts import from_nested
from util import small_synth, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _pair(K, alpha, eta, V, csr, lam0, arith=orc.ARITH_LITERAL):
    o = orc.OracleLDA(K, alpha, eta, V, csr, lam0, arith=arith, nthreads=8)
    g = mmsig.LDA(K, alpha, eta, csr, V=V, lambda0=lam0)
    return o, g


def _check(o, g, ll_o, ll_g, tol=TOL):
    s = g.state()
    assert rel_err(s["lam"], o.lam) <= tol
    assert rel_err(s["gamma"], o.gamma) <= tol
    assert rel_err(s["beta"], o.beta) <= tol
    assert rel_err(s["theta"], o.theta) <= tol
    assert np.max(np.abs(s["Elnbeta"] - o.Elnbeta)) <= tol * 50      # differences of O(10) digammas
    assert np.max(np.abs(s["Elntheta"] - o.Elntheta)) <= tol * 50
    assert abs(ll_g - ll_o) <= tol * abs(ll_o)


def test_lda_toy(golden):
    t = golden["lda_toy"]
    csr = from_nested([[x] for x in t["X"]], 1)[0]
    lam0 = np.array([3., 50, 17, 99])
    o, g = _pair(t["K"], t["alpha"], t["eta"], t["V"], csr, lam0)
    for _ in range(3):
        ll_o, ll_g = o.iterate(), g.iterate()
        _check(o, g, ll_o, ll_g)
    np.testing.assert_allclose(g.phi(), o.phi, rtol=1e-12)
    eo, to = o.elbo()
    eg, tg = g.calculate_elbo()
    np.testing.assert_allclose(tg, to, rtol=1e-11, atol=1e-12)
    assert abs(eg - eo) <= 1e-11 * abs(eo)
    g.close()


@pytest.mark.parametrize("arith", [orc.ARITH_LITERAL, orc.ARITH_DET])
@pytest.mark.parametrize("K,V,D", [(20, 96, 3000), (7, 48, 500), (1, 5, 40), (32, 12, 200)])
def test_lda_synthetic(K, V, D, arith):
    csr = small_synth(D, [K], [V], empty_frac=0.05)[0]
    lam0 = mmsig.synth.init_lda_lambda(K, V)
    o, g = _pair(K, 0.1, 0.1, V, csr, lam0, arith=arith)
    for _ in range(4):
        ll_o, ll_g = o.iterate(), g.iterate()
        _check(o, g, ll_o, ll_g)
    eo, to = o.elbo()
    eg, tg = g.calculate_elbo()
    assert abs(eg - eo) <= 1e-11 * abs(eo), (tg, to)
    g.close()


def test_lda_fit_config2_shape(brca):
    """LDA(20, 0.1, 0.1) (config 2 shape) on the brca-eu SNV counts, fixed 25 iterations + tol stop."""
    csr = brca[0]
    K, V = 20, 96
    lam0 = mmsig.synth.init_lda_lambda(K, V)
    o, g = _pair(K, 0.1, 0.1, V, csr, lam0)
    ho = o.fit(maxiter=25, tol=1e-7)
    hg = g.fit(maxiter=25, tol=1e-7, verbose=False)
    assert hg.shape == ho.shape and g.converged == o.converged
    assert rel_err(hg, ho) <= 1e-11
    eo, _ = o.elbo()
    assert abs(g.elbo - eo) <= 1e-8 * abs(eo)
    assert rel_err(g.beta, o.beta) <= 1e-9
    g.close()


def test_lda_fit_host_equals_the_four_calls(brca):
    """mmsig_lda_fit_host == set_data + set_state + fit + get_state, bit for bit (a handle planned for another corpus)."""
    K, V = 20, 96
    csr = brca[0]
    lam0 = mmsig.synth.init_lda_lambda(K, V)
    a = mmsig.LDA(K, 0.1, 0.1, csr, V=V, lambda0=lam0)
    ha = a.fit(maxiter=14, tol=1e-7, verbose=False)
    sa = a.state()
    other = small_synth(300, [K], [V])[0]
    b = mmsig.LDA(K, 0.1, 0.1, other, V=V, lambda0=lam0)
    hb, sb = b.fit_host(csr, lam0, maxiter=14, tol=1e-7)
    assert np.array_equal(ha, hb) and a.converged == b.converged
    for k in ("lam", "Elnbeta", "beta", "gamma", "Elntheta", "theta"):
        assert np.array_equal(np.asarray(sa[k]).reshape(-1), np.asarray(sb[k]).reshape(-1)), k
    a.close(); b.close()

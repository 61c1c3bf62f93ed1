"""Next rows of SURVEY 8(f)-1 on the same kernels with the M-step (partly) frozen:
fit_heldout, transform (unsmoothed θ), predict_modality_η for MMCTM; transform, fit_heldout for LDA."""
import numpy as np
import pytest

import orc
import mmsig
from util import oracle_mmctm, small_synth, rel_err

pytestmark = pytest.mark.gpu
F = orc


def _fit_pair(K, V, D):
    counts = small_synth(D, K, V)
    alpha = [0.1] * len(K)
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, alpha, V, counts, g0)
    g = mmsig.MMCTM(K, alpha, counts, V=V, gamma0=g0)
    for _ in range(4):
        o.iterate(); g.iterate()
    return o, g, alpha


def _oracle_child(o, K, alpha, V, counts, gamma, copy_inv=True, copy_gauss=True):
    c = oracle_mmctm(K, alpha, V, counts, gamma)
    if copy_gauss:
        c.mu[:] = o.mu; c.Sigma[:] = o.Sigma
        if copy_inv:
            c.invSigma[:] = o.invSigma
    c.gamma[:] = gamma
    c.L.orc_mmctm_update_Elnphi(c.p)
    c.phi[:] = o.phi
    return c


def test_mmctm_fit_heldout():
    K, V = [10, 8, 6], [96, 32, 83]
    o, g, alpha = _fit_pair(K, V, 600)
    held = small_synth(250, K, V, seed=99, empty_frac=0.05)
    oc = _oracle_child(o, K, alpha, V, held, o.gamma.copy())
    gh = g.fit_heldout(held, maxiter=13)
    flags = F.FLAG_FREEZE_TOPICS | F.FLAG_FREEZE_MU
    ho = []
    for _ in range(len(gh.ll_history)):
        ho.append(oc.iterate_flags(flags))
    assert np.array_equal(gh.ll_history, np.asarray(ho))
    s = gh.state()
    assert np.array_equal(s["lam"], oc.lam) and np.array_equal(s["nu"], oc.nu) and np.array_equal(s["props"], oc.props)
    assert np.array_equal(s["phi"], o.phi) and np.array_equal(s["mu"], o.mu)          # frozen
    gh.close(); g.close()


@pytest.mark.parametrize("fit_gaussian", [False, True])
def test_mmctm_transform(fit_gaussian):
    K, V = [7, 7], [96, 32]
    o, g, alpha = _fit_pair(K, V, 500)
    newc = small_synth(200, K, V, seed=123)
    gt = g.transform(newc, maxiter=5, tol=1e-12, fit_gaussian=fit_gaussian, rng=np.random.default_rng(3))
    g0 = np.random.default_rng(3).integers(1, 101, size=g.G).astype(float)
    oc = _oracle_child(o, K, alpha, V, newc, g0, copy_inv=False, copy_gauss=not fit_gaussian)
    flags = F.FLAG_FREEZE_TOPICS | F.FLAG_UNSMOOTHED | (F.FLAG_UPDATE_SIGMA if fit_gaussian else F.FLAG_FREEZE_MU)
    ho = np.asarray([oc.iterate_flags(flags) for _ in range(5)])
    assert np.array_equal(gt.ll_history, ho)
    s = gt.state()
    assert np.array_equal(s["lam"], oc.lam) and np.array_equal(s["Sigma"], oc.Sigma) and np.array_equal(s["props"], oc.props)
    np.testing.assert_allclose(gt.theta(0), oc.theta(0), rtol=1e-13)
    if fit_gaussian:
        assert not np.array_equal(s["Sigma"], o.Sigma)          # test/mmctm.jl:400-403
    else:
        assert np.array_equal(s["Sigma"], o.Sigma)              # test/mmctm.jl:394-398
    gt.close(); g.close()


def test_mmctm_predict_modality_eta():
    K, V = [5, 4, 3], [40, 20, 30]
    o, g, alpha = _fit_pair(K, V, 300)
    obs_counts = small_synth(120, K, V, seed=77)
    m = 1
    obs = [obs_counts[0], obs_counts[2]]
    eta = g.predict_modality_eta(obs, m, maxiter=6)
    ob = np.r_[0:5, 9:12]; un = np.r_[5:9]
    go = np.cumsum([0] + [k * v for k, v in zip(K, V)])
    g_obs = np.concatenate([o.gamma[go[0]:go[1]], o.gamma[go[2]:go[3]]])
    oc = oracle_mmctm([5, 3], [0.1, 0.1], [40, 30], obs, g_obs)
    oc.mu[:] = o.mu[ob]; oc.Sigma[:] = o.Sigma[np.ix_(ob, ob)]; oc.invSigma[:] = o.invSigma[np.ix_(ob, ob)]
    for _ in range(6):
        oc.iterate_flags(F.FLAG_FREEZE_TOPICS | F.FLAG_FREEZE_MU)
    ref = o.mu[un] + (oc.lam - o.mu[ob]) @ (o.Sigma[np.ix_(un, ob)] @ o.invSigma[np.ix_(ob, ob)]).T
    assert eta.shape == (120, 4)
    np.testing.assert_array_equal(eta, ref)
    g.close()


def test_lda_transform_and_heldout():
    K, V = 12, 60
    csr = small_synth(800, [K], [V])[0]
    lam0 = mmsig.synth.init_lda_lambda(K, V)
    o = orc.OracleLDA(K, 0.1, 0.1, V, csr, lam0, nthreads=8)
    g = mmsig.LDA(K, 0.1, 0.1, csr, V=V, lambda0=lam0)
    for _ in range(5):
        o.iterate(); g.iterate()
    new = small_synth(300, [K], [V], seed=5)[0]
    # transform
    th = g.transform(new, maxiter=6, tol=0.0)
    oc = orc.OracleLDA(K, 0.1, 0.1, V, new, np.ones(K * V), nthreads=8)
    oc.beta[:] = o.beta
    for _ in range(6):
        oc.iterate_flags(F.FLAG_FREEZE_TOPICS | F.FLAG_UNSMOOTHED)
    assert rel_err(th, oc.theta) <= 1e-12
    # fit_heldout
    gh = g.fit_heldout(new, maxiter=6)
    oh = orc.OracleLDA(K, 0.1, 0.1, V, new, o.lam.ravel().copy(), nthreads=8)
    oh.beta[:] = o.beta
    ho = np.asarray([oh.iterate_flags(F.FLAG_FREEZE_TOPICS) for _ in range(6)])
    assert rel_err(gh.ll_history, ho) <= 1e-12
    assert rel_err(gh.gamma, oh.gamma) <= 1e-12
    eo = oh.elbo()[0]
    assert abs(gh.elbo - eo) <= 1e-11 * abs(eo)
    gh.close(); g.close()

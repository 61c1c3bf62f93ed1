"""IMMCTM (reference src/IMMCTM.jl) on the MMCTM's kernels with the feature-table M-step
(k_imstep1): against the oracle's pinned arithmetic, bit-exact like the MMCTM."""
import numpy as np
import pytest

import orc
import mmsig
from mmsig.counts import make_count_csr

pytestmark = pytest.mark.gpu


def _grid_features(*sizes):
    """All combinations of feature values (e.g. mutation type x context), one term per row."""
    g = np.stack(np.meshgrid(*[np.arange(s) for s in sizes], indexing="ij"), -1).reshape(-1, len(sizes))
    return np.ascontiguousarray(g, dtype=np.int32)


def _case(seed, D, K, feats, autoalpha=False, iters=3):
    rng = np.random.default_rng(seed)
    counts = [make_count_csr(rng.poisson(2.0, size=(f.shape[0], D))) for f in feats]
    J = [[int(f[:, i].max()) + 1 for i in range(f.shape[1])] for f in feats]
    T = sum(k * sum(j) for k, j in zip(K, J))
    g0 = rng.integers(1, 101, T).astype(float)
    alpha = [0.1 + 0.05 * m for m in range(len(K))]
    o = orc.OracleIMMCTM(K, alpha, feats, counts, g0, arith=orc.ARITH_DET, nthreads=8)
    g = mmsig.IMMCTM(K, alpha, feats, counts, gammaf0=g0)
    for _ in range(iters):
        ll_o = o.iterate(autoalpha=autoalpha)
        ll_g = g.iterate(flags=mmsig.capi.FLAG_UPDATE_SIGMA | (mmsig.capi.FLAG_AUTO_ALPHA if autoalpha else 0))
        s, t = g.state(), g.tables()
        for k in ("lam", "nu", "zeta", "mu", "Sigma", "invSigma", "Elnphi", "phi", "props"):
            assert np.array_equal(s[k], getattr(o, k)), k
        assert np.array_equal(t["gammaf"], o.gammaf) and np.array_equal(t["Elnphif"], o.Elnphif)
        assert np.array_equal(ll_g, ll_o)
        if autoalpha:
            np.testing.assert_allclose(t["alphaf"], o.alphaf, rtol=1e-12)
    eo, eg = o.elbo(), g.calculate_elbo()
    assert abs(eg[0] - eo[0]) <= 1e-12 * abs(eo[0]), (eg, eo)
    np.testing.assert_allclose(eg[1], eo[1], rtol=1e-10, atol=1e-6)
    g.close()


def test_immctm_two_modalities():
    _case(1, 300, [3, 2], [_grid_features(6, 4, 4), _grid_features(5)])       # SNV-like 96 terms as 6 x 4 x 4, plus a flat modality


def test_immctm_three_modalities_and_a_one_feature_model():
    _case(2, 150, [4, 3, 2], [_grid_features(3, 2), _grid_features(2, 2, 2), _grid_features(7)], iters=2)
    _case(3, 97, [2], [_grid_features(9)], iters=2)


def test_immctm_auto_alpha():
    _case(4, 200, [3, 2], [_grid_features(4, 3), _grid_features(2, 5)], autoalpha=True, iters=2)


def test_immctm_refuses_what_it_has_not():
    feats = [_grid_features(3, 2)]
    rng = np.random.default_rng(5)
    counts = [make_count_csr(rng.poisson(2.0, size=(6, 40)))]
    g = mmsig.IMMCTM([2], [0.1], feats, counts, rng=rng)
    with pytest.raises(mmsig.capi.MmsigError):
        g.iterate(flags=mmsig.capi.FLAG_UNSMOOTHED | mmsig.capi.FLAG_FREEZE_TOPICS)
    g.close()


def test_immctm_restarts():
    """Independent restarts (README.md:42) of the IMMCTM: per-restart ELBO, LL and the best state."""
    rng = np.random.default_rng(8)
    feats = [_grid_features(4, 3), _grid_features(6)]
    K, alpha = [3, 2], [0.1, 0.1]
    counts = [make_count_csr(rng.poisson(2.0, size=(f.shape[0], 150))) for f in feats]
    T = 3 * 7 + 2 * 6
    g0s = rng.integers(1, 101, size=(3, T)).astype(float)
    g = mmsig.IMMCTM(K, alpha, feats, counts, gammaf0=g0s[0])
    elbo, ll, nit, best = g.fit_restarts(g0s, maxiter=4)
    ref = []
    for r in range(3):
        o = orc.OracleIMMCTM(K, alpha, feats, counts, g0s[r], arith=orc.ARITH_DET, nthreads=8)
        h = o.fit(maxiter=4)
        ref.append((o.elbo()[0], h[-1], o.gammaf.copy(), o.lam.copy()))
        assert np.array_equal(ll[r], h[-1])
        assert abs(elbo[r] - ref[r][0]) <= 1e-12 * abs(ref[r][0])
    assert best == int(np.argmax([x[0] for x in ref])) and list(nit) == [4, 4, 4]
    assert np.array_equal(g.tables()["gammaf"], ref[best][2]) and np.array_equal(g.state()["lam"], ref[best][3])
    g.close()


def test_immctm_fit_heldout():
    """fit_heldout (src/IMMCTM.jl:547-579): frozen feature tables and Gaussian prior, E-step + LL only."""
    rng = np.random.default_rng(6)
    feats = [_grid_features(4, 3), _grid_features(6)]
    K, alpha = [3, 2], [0.1, 0.2]
    tr = [make_count_csr(rng.poisson(2.0, size=(f.shape[0], 200))) for f in feats]
    ho = [make_count_csr(rng.poisson(2.0, size=(f.shape[0], 77))) for f in feats]
    T = sum(k * sum(int(f[:, i].max()) + 1 for i in range(f.shape[1])) for k, f in zip(K, feats))
    g0 = rng.integers(1, 101, T).astype(float)
    o = orc.OracleIMMCTM(K, alpha, feats, tr, g0, arith=orc.ARITH_DET, nthreads=8)
    g = mmsig.IMMCTM(K, alpha, feats, tr, gammaf0=g0)
    for _ in range(3):
        o.iterate(); g.iterate()
    oh = orc.OracleIMMCTM(K, alpha, feats, ho, o.gammaf.copy(), arith=orc.ARITH_DET, nthreads=8)
    oh.mu[:] = o.mu; oh.Sigma[:] = o.Sigma; oh.invSigma[:] = o.invSigma
    ll_o = [oh.iterate_flags(orc.FLAG_FREEZE_TOPICS | orc.FLAG_FREEZE_MU) for _ in range(4)]
    gh = g.fit_heldout(ho, maxiter=4)
    assert np.array_equal(gh.ll_history, np.asarray(ll_o))
    s = gh.state()
    for k in ("lam", "nu", "zeta", "props", "phi", "Elnphi"):
        assert np.array_equal(s[k], getattr(oh, k)), k
    assert np.array_equal(gh.tables()["gammaf"], o.gammaf)
    g.close(); gh.close()


def test_immctm_predict_modality_eta():
    """predict_modality_η (src/IMMCTM.jl:581-627): the observed modalities' sub-model with frozen tables."""
    rng = np.random.default_rng(7)
    feats = [_grid_features(4, 3), _grid_features(2, 2, 2), _grid_features(5)]
    K, alpha = [3, 2, 2], [0.1, 0.1, 0.1]
    tr = [make_count_csr(rng.poisson(2.0, size=(f.shape[0], 250))) for f in feats]
    J = [[int(f[:, i].max()) + 1 for i in range(f.shape[1])] for f in feats]
    ts = np.cumsum([0] + [k * sum(j) for k, j in zip(K, J)])
    g0 = rng.integers(1, 101, ts[-1]).astype(float)
    o = orc.OracleIMMCTM(K, alpha, feats, tr, g0, arith=orc.ARITH_DET, nthreads=8)
    g = mmsig.IMMCTM(K, alpha, feats, tr, gammaf0=g0)
    for _ in range(3):
        o.iterate(); g.iterate()
    m = 1
    obs = [make_count_csr(rng.poisson(2.0, size=(feats[i].shape[0], 60))) for i in (0, 2)]
    eta = g.predict_modality_eta(obs, m, maxiter=5)
    ob, un = np.r_[0:3, 5:7], np.r_[3:5]
    g_obs = np.concatenate([o.gammaf[ts[0]:ts[1]], o.gammaf[ts[2]:ts[3]]])
    oc = orc.OracleIMMCTM([3, 2], [0.1, 0.1], [feats[0], feats[2]], obs, g_obs, arith=orc.ARITH_DET, nthreads=8)
    oc.mu[:] = o.mu[ob]; oc.Sigma[:] = o.Sigma[np.ix_(ob, ob)]; oc.invSigma[:] = o.invSigma[np.ix_(ob, ob)]
    for _ in range(5):
        oc.iterate_flags(orc.FLAG_FREEZE_TOPICS | orc.FLAG_FREEZE_MU)
    ref = o.mu[un] + (oc.lam - o.mu[ob]) @ (o.Sigma[np.ix_(un, ob)] @ o.invSigma[np.ix_(ob, ob)]).T
    assert eta.shape == (60, 2)
    np.testing.assert_array_equal(eta, ref)
    g.close()

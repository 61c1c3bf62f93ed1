"""The oracle against a second, vectorised numpy statement: the ELBOs (seven terms each)
(MMCTM: reference src/MMCTM.jl:271-382; LDA: src/LDA.jl:114-172), evaluated on the oracle's own
state after a few iterations.  The reference's tests only check the sign of the ELBO; the CUDA path
is held to the oracle's value (1e-12), so the oracle's transcription is cross-checked here."""
import numpy as np
import pytest
from scipy.special import gammaln

import orc
import mmsig
from util import small_synth

ARITHS = [orc.ARITH_LITERAL, orc.ARITH_DET]


def _expand(rowptr):
    return np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))


def mmctm_elbo_numpy(o, counts):
    K, V, M, D, MK = [int(k) for k in o.K], [int(v) for v in o.V], o.M, o.D, o.MK
    koff = np.concatenate([[0], np.cumsum(K)])
    goff = np.concatenate([[0], np.cumsum(np.array(K) * np.array(V))])
    lam, nu, zeta, mu, S, N = o.lam, o.nu, o.zeta, o.mu, o.invSigma, o.N()
    t = np.zeros(7)
    sumtheta = np.zeros((D, MK))
    for m in range(M):
        a = float(o.alpha[m])
        E = o.Elnphi[goff[m]:goff[m + 1]].reshape(K[m], V[m])
        g = o.gamma[goff[m]:goff[m + 1]].reshape(K[m], V[m])
        t[0] += -K[m] * (V[m] * gammaln(a) - gammaln(V[m] * a)) + (a - 1) * E.sum()              # :271-284
        t[4] += -(gammaln(g).sum(1) - gammaln(g.sum(1))).sum() + ((g - 1) * E).sum()              # :338-350
        rp, term, cnt = counts[m]
        d_of = _expand(rp)
        th = o.theta(m)                                                                          # (nnz, K_m)
        nth = th * cnt[:, None]
        np.add.at(sumtheta[:, koff[m]:koff[m + 1]], d_of, nth)                                   # :110-117
        t[3] += (nth * E[:, term].T).sum()                                                       # :318-336
        with np.errstate(divide="ignore", invalid="ignore"):
            tl = np.where(th > 0, th * np.log(th), 0.0)                                          # log(θ^θ), 0^0 = 1
        t[6] += (cnt[:, None] * tl).sum()                                                        # :360-370
    diff = lam - mu
    _, logdet = np.linalg.slogdet(S)
    quad = np.einsum("di,ij,dj->d", diff, S, diff)
    t[1] = 0.5 * (D * (logdet - MK * np.log(2 * np.pi)) - (nu @ np.diag(S)).sum() - quad.sum())   # :286-300
    Ee = np.exp(lam + 0.5 * nu)
    Ndz = np.repeat(N / zeta, K, axis=1)
    t[2] = (lam * sumtheta).sum() - (Ndz * Ee).sum() + N.sum() - (N * np.log(zeta)).sum()        # :302-316
    t[5] = -0.5 * (np.log(nu).sum() + D * MK * (np.log(2 * np.pi) + 1))                           # :352-358
    return t[0] + t[1] + t[2] + t[3] - t[4] - t[5] - t[6], t


@pytest.mark.parametrize("arith", ARITHS)
def test_mmctm_elbo(arith):
    K, V, D = [3, 4], [12, 9], 150
    counts = small_synth(D, K, V, empty_frac=0.05)
    o = orc.OracleMMCTM(K, [0.1, 0.3], V, counts, mmsig.synth.init_gamma(K, V), arith=arith, nthreads=4)
    for _ in range(3):
        o.iterate()
    e, t = o.elbo()
    e2, t2 = mmctm_elbo_numpy(o, counts)
    np.testing.assert_allclose(t, t2, rtol=1e-10, atol=1e-9 * abs(e))
    assert abs(e - e2) <= 1e-10 * abs(e)


def lda_elbo_numpy(o, counts, alpha, eta):
    K, V, D = o.K, o.V, o.D
    rp, term, cnt = counts
    d_of = _expand(rp)
    lam, Eb, g, Et, ph = o.lam, o.Elnbeta, o.gamma, o.Elntheta, o.phi
    t = np.zeros(7)
    t[0] = K * (gammaln(V * eta) - V * gammaln(eta)) + (eta - 1) * Eb.sum()                 # :114-118
    t[1] = D * (gammaln(K * alpha) - K * gammaln(alpha)) + (alpha - 1) * Et.sum()           # :120-124
    t[2] = (ph * Et[d_of] * cnt[:, None]).sum()                                                  # :126-132
    t[3] = (ph * Eb[:, term].T * cnt[:, None]).sum()                                             # :134-140
    t[4] = gammaln(lam).sum() - gammaln(lam.sum(1)).sum() - ((lam - 1) * Eb).sum()                # :142-146
    t[5] = gammaln(g).sum() - gammaln(g.sum(1)).sum() - ((g - 1) * Et).sum()                      # :148-152
    with np.errstate(divide="ignore", invalid="ignore"):
        t[6] = np.where(ph > 0, ph * np.log(ph), 0.0).sum()                                      # :154-160, NOT count weighted
    return t[0] + t[1] + t[2] + t[3] - t[4] - t[5] - t[6], t


@pytest.mark.parametrize("arith", ARITHS)
def test_lda_elbo(arith):
    K, V, D = 5, 20, 300
    csr = small_synth(D, [K], [V], empty_frac=0.05)[0]
    o = orc.OracleLDA(K, 0.1, 0.2, V, csr, mmsig.synth.init_lda_lambda(K, V), arith=arith, nthreads=4)
    for _ in range(3):
        o.iterate()
    e, t = o.elbo()
    e2, t2 = lda_elbo_numpy(o, csr, 0.1, 0.2)
    np.testing.assert_allclose(t, t2, rtol=1e-10, atol=1e-10 * abs(e))
    assert abs(e - e2) <= 1e-10 * abs(e)


def _grid_features(shape):
    return np.stack(np.meshgrid(*[range(s) for s in shape], indexing="ij"), -1).reshape(-1, len(shape))


@pytest.mark.parametrize("arith", ARITHS)
def test_immctm_elbo_table_terms(arith):
    """IMMCTM: ElnPϕ (src/IMMCTM.jl:247-261) and ElnQϕ (:314-329) run over the feature tables; the other
    five terms are the MMCTM's over the composite Elnϕ (checked through mmctm_elbo_numpy)."""
    K, shapes, D = [3, 2], [(3, 4), (2, 3)], 120
    feats = [_grid_features(s) for s in shapes]
    V = [f.shape[0] for f in feats]
    counts = small_synth(D, K, V)
    alphaf = [[0.1, 0.4], [0.2, 0.3]]
    T = sum(k * sum(s) for k, s in zip(K, shapes))
    o = orc.OracleIMMCTM(K, alphaf, feats, counts, np.random.default_rng(4).integers(1, 101, T).astype(float),
                         arith=arith, nthreads=4)
    for _ in range(3):
        o.iterate()
    e, t = o.elbo()
    _, t2 = mmctm_elbo_numpy(o, counts)
    np.testing.assert_allclose(t[[1, 2, 3, 5, 6]], t2[[1, 2, 3, 5, 6]], rtol=1e-10, atol=1e-9 * abs(e))
    p = q = 0.0
    for m in range(2):
        for k in range(K[m]):
            for i in range(len(shapes[m])):
                a, J = alphaf[m][i], shapes[m][i]
                E, g = o.table(o.Elnphif, m, k, i), o.table(o.gammaf, m, k, i)
                p += -(J * gammaln(a) - gammaln(J * a)) + (a - 1) * E.sum()
                q += -(gammaln(g).sum() - gammaln(g.sum())) + ((g - 1) * E).sum()
    np.testing.assert_allclose([t[0], t[4]], [p, q], rtol=1e-11)
    assert abs(e - (p + t2[1] + t2[2] + t2[3] - q - t2[5] - t2[6])) <= 1e-10 * abs(e)


@pytest.mark.parametrize("arith", ARITHS)
def test_ilda_elbo_table_terms(arith):
    """ILDA: ElnPβ per feature (src/ILDA.jl:131-140); ElnQβ AS WRITTEN at :174-181 — the `=` inside the loop
    keeps only the last feature's term."""
    K, shape, D = 4, (3, 5), 200
    feat = _grid_features(shape)
    V = feat.shape[0]
    csr = small_synth(D, [K], [V])[0]
    eta = [0.2, 0.6]
    o = orc.OracleILDA(K, 0.1, eta, feat, csr, np.random.default_rng(6).integers(1, 101, K * sum(shape)).astype(float),
                       arith=arith, nthreads=4)
    for _ in range(3):
        o.iterate()
    e, t = o.elbo()
    _, t2 = lda_elbo_numpy(o, csr, 0.1, eta[0])
    np.testing.assert_allclose(t[[1, 2, 3, 5, 6]], t2[[1, 2, 3, 5, 6]], rtol=1e-10, atol=1e-10 * abs(e))
    p, q = 0.0, None
    for i, J in enumerate(shape):
        lam = np.stack([o.table(o.lambdaf, k, i) for k in range(K)])          # (K, J_i)
        Eb = np.stack([o.table(o.Elnbetaf, k, i) for k in range(K)])
        p += K * (gammaln(J * eta[i]) - J * gammaln(eta[i])) + (eta[i] - 1) * Eb.sum()
        q = gammaln(lam).sum() - gammaln(lam.sum(1)).sum() - ((lam - 1) * Eb).sum()
    np.testing.assert_allclose([t[0], t[4]], [p, q], rtol=1e-11)


@pytest.mark.parametrize("arith", ARITHS)
def test_loglikelihoods_numpy(arith):
    """calculate_loglikelihoods (src/MMCTM.jl:384-448) and LDA's (src/LDA.jl:174-188) on data with empty rows:
    ll_m = Σ_d Σ_w n log(props_d · ϕ[:, v]) / Σ_d N_dm."""
    K, V, D = [3, 4], [12, 9], 150
    counts = small_synth(D, K, V, empty_frac=0.1)
    o = orc.OracleMMCTM(K, [0.1, 0.3], V, counts, mmsig.synth.init_gamma(K, V), arith=arith, nthreads=4)
    ll = None
    for _ in range(2):
        ll = o.iterate()
    koff = np.concatenate([[0], np.cumsum(K)])
    goff = np.concatenate([[0], np.cumsum(np.array(K) * np.array(V))])
    for m in range(2):
        rp, term, cnt = counts[m]
        assert np.any(np.diff(rp) == 0)                                        # the empty rows are there
        d_of = _expand(rp)
        ph = o.phi[goff[m]:goff[m + 1]].reshape(K[m], V[m])
        pr = o.props[:, koff[m]:koff[m + 1]]
        pw = np.einsum("wk,kw->w", pr[d_of], ph[:, term])
        np.testing.assert_allclose(ll[m], (cnt * np.log(pw)).sum() / cnt.sum(), rtol=1e-12)
    csr = small_synth(300, [5], [20], empty_frac=0.1)[0]
    l = orc.OracleLDA(5, 0.1, 0.2, 20, csr, mmsig.synth.init_lda_lambda(5, 20), arith=arith, nthreads=4)
    for _ in range(2):
        v = l.iterate()
    rp, term, cnt = csr
    pw = np.einsum("wk,kw->w", l.theta[_expand(rp)], l.beta[:, term])
    np.testing.assert_allclose(v, (cnt * np.log(pw)).sum() / cnt.sum(), rtol=1e-12)


@pytest.mark.parametrize("arith", ARITHS)
def test_mstep_numpy(arith):
    """The M-step of one iteration (src/MMCTM.jl:200-250, :145-154) recomputed with numpy from the oracle's own
    per-sample results (λ, ν, θ): μ, Σ, invΣ, γ, Elnϕ, ϕ, props."""
    from scipy.special import digamma
    K, V, D = [3, 4], [12, 9], 200
    counts = small_synth(D, K, V, empty_frac=0.05)
    al = [0.1, 0.3]
    o = orc.OracleMMCTM(K, al, V, counts, mmsig.synth.init_gamma(K, V), arith=arith, nthreads=4)
    for _ in range(2):
        o.iterate()
    lam, nu = o.lam, o.nu
    mu = lam.mean(0)
    np.testing.assert_allclose(o.mu, mu, rtol=1e-12, atol=1e-14)
    diff = lam - mu
    Sig = (np.diag(nu.sum(0)) + diff.T @ diff) / D
    np.testing.assert_allclose(o.Sigma, Sig, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(o.invSigma, np.linalg.inv(Sig), rtol=1e-9, atol=1e-11)
    koff = np.concatenate([[0], np.cumsum(K)])
    goff = np.concatenate([[0], np.cumsum(np.array(K) * np.array(V))])
    for m in range(2):
        rp, term, cnt = counts[m]
        g = np.full((K[m], V[m]), al[m])
        np.add.at(g.T, term, o.theta(m) * cnt[:, None])
        og = o.gamma[goff[m]:goff[m + 1]].reshape(K[m], V[m])
        np.testing.assert_allclose(og, g, rtol=1e-11)
        np.testing.assert_allclose(o.Elnphi[goff[m]:goff[m + 1]].reshape(K[m], V[m]),
                                   digamma(og) - digamma(og.sum(1))[:, None], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(o.phi[goff[m]:goff[m + 1]].reshape(K[m], V[m]), og / og.sum(1)[:, None], rtol=1e-13)
        e = np.exp(lam[:, koff[m]:koff[m + 1]])
        np.testing.assert_allclose(o.props[:, koff[m]:koff[m + 1]], e / e.sum(1)[:, None], rtol=1e-12)


@pytest.mark.parametrize("arith", ARITHS)
def test_feature_table_msteps_numpy(arith):
    """update_γ! of the IMMCTM (src/IMMCTM.jl:197-221) and update_λ! of the ILDA (src/ILDA.jl:105-125) on random data:
    table_k,i,j = prior_i + Σ over the nonzeros whose term carries value j of feature i of n·(θ or ϕ)_k."""
    K, shapes, D = [3, 2], [(3, 4), (2, 3)], 150
    feats = [_grid_features(s) for s in shapes]
    V = [f.shape[0] for f in feats]
    counts = small_synth(D, K, V, empty_frac=0.05)
    alphaf = [[0.1, 0.4], [0.2, 0.3]]
    T = sum(k * sum(s) for k, s in zip(K, shapes))
    o = orc.OracleIMMCTM(K, alphaf, feats, counts, np.random.default_rng(4).integers(1, 101, T).astype(float),
                         arith=arith, nthreads=4)
    for _ in range(2):
        o.iterate()
    for m in range(2):
        rp, term, cnt = counts[m]
        nth = o.theta(m) * cnt[:, None]
        for k in range(K[m]):
            for i, J in enumerate(shapes[m]):
                want = np.full(J, alphaf[m][i])
                np.add.at(want, feats[m][term, i], nth[:, k])
                np.testing.assert_allclose(o.table(o.gammaf, m, k, i), want, rtol=1e-11)
    K1, shape = 4, (3, 5)
    feat = _grid_features(shape)
    csr = small_synth(200, [K1], [feat.shape[0]], empty_frac=0.05)[0]
    eta = [0.2, 0.6]
    l = orc.OracleILDA(K1, 0.1, eta, feat, csr, np.random.default_rng(6).integers(1, 101, K1 * sum(shape)).astype(float),
                       arith=arith, nthreads=4)
    for _ in range(2):
        l.iterate()
    rp, term, cnt = csr
    nph = l.phi * cnt[:, None]
    for k in range(K1):
        for i, J in enumerate(shape):
            want = np.full(J, eta[i])
            np.add.at(want, feat[term, i], nph[:, k])
            np.testing.assert_allclose(l.table(l.lambdaf, k, i), want, rtol=1e-11)

"""The oracle (LITERAL and DET arithmetic) against the known answers of the reference's own
tests (tests/golden/reference_known_answers.json <- /root/reference/test/*.jl)."""
import ctypes as C

import numpy as np
import pytest

import orc
from mmsig.counts import from_nested

RTOL = 1e-13          # the reference's own `≈` is rtol ~1.5e-8; the oracle is held far tighter


def _toy(golden, arith, gamma0=None):
    t = golden["mmctm_toy"]
    counts = from_nested(t["X"], 2)
    g0 = np.ones(2 * 4 + 3 * 4) if gamma0 is None else np.asarray(gamma0, float)
    return orc.OracleMMCTM(t["K"], t["alpha"], [4, 4], counts, g0, arith=arith)


ARITHS = [orc.ARITH_LITERAL, orc.ARITH_DET]


@pytest.mark.parametrize("arith", ARITHS)
def test_ctor(golden, arith):
    m = _toy(golden, arith)
    g = golden["ctor"]
    assert m.D == g["D"] and m.M == g["M"] and m.MK == g["MK"]
    assert m.N().tolist() == g["N"]
    assert np.all(m.nu == 1.0) and np.all(m.lam == 0.0)
    np.testing.assert_allclose(m.theta(0).sum(axis=1), 1.0, rtol=1e-15)
    from mmsig.counts import infer_V
    assert infer_V(from_nested(golden["mmctm_toy"]["X"], 2)) == g["V"]


@pytest.mark.parametrize("arith", ARITHS)
def test_Ndivzeta(golden, arith):
    m = _toy(golden, arith)
    m.zeta[:] = np.asarray(golden["Ndivzeta"]["zeta"], float)
    out = np.zeros(5)
    m.L.orc_mmctm_calc_Ndivzeta(m.p, 0, orc._dp(out))
    np.testing.assert_allclose(out, golden["Ndivzeta"]["expected_d1"], rtol=RTOL)


def _set_theta(m, mod, d, th):
    """th: K x W as in the reference; oracle stores nnz x K."""
    rp = m._keep[mod][0]
    m.theta(mod)[rp[d]:rp[d + 1], :] = np.asarray(th, float).T


def test_sumtheta(golden):
    m = _toy(golden, orc.ARITH_LITERAL)
    th = golden["sumtheta"]["theta_d1"]
    _set_theta(m, 0, 0, th[0]); _set_theta(m, 1, 0, th[1])
    out = np.zeros(5)
    m.L.orc_mmctm_calc_sumtheta(m.p, 0, orc._dp(out))
    np.testing.assert_allclose(out, golden["sumtheta"]["expected_d1"], rtol=RTOL)


def test_sumtheta_det_product_form_equals_definition(golden):
    """The pinned specification evaluates calculate_sumθ (src/MMCTM.jl:110-117) in the product form
    exp(λ_k) Σ_w E_kv n_w / Z_w on the E-step's own intermediates instead of on the stored θ; it must
    equal Σ_w n_w θ_kw of the θ the same E-step stored, to rounding."""
    m = _toy(golden, orc.ARITH_DET)
    for d in range(m.D):
        m.L.orc_mmctm_update_theta(m.p, d)
        out = np.zeros(m.MK)
        m.L.orc_mmctm_calc_sumtheta(m.p, d, orc._dp(out))
        ref, off = np.zeros(m.MK), 0
        for mod in range(m.M):
            rp, _, cnt = m._keep[mod]
            th = m.theta(mod)[rp[d]:rp[d + 1], :]
            K = th.shape[1]
            ref[off:off + K] = (th * np.asarray(cnt[rp[d]:rp[d + 1]], float)[:, None]).sum(axis=0)
            off += K
        np.testing.assert_allclose(out, ref, rtol=1e-13)


@pytest.mark.parametrize("arith", ARITHS)
def test_lambda_objective(golden, arith):
    g = golden["lambda_objective"]
    L = orc.lib()
    lam, nu, mu = (np.asarray(g[k], float) for k in ("lambda", "nu", "mu"))
    c = np.array([13 / g["zeta"][0]] * 2 + [7 / g["zeta"][1]] * 3)
    st = np.asarray(golden["sumtheta"]["expected_d1"])
    S = np.eye(5)
    grad = np.zeros(5)
    v = L.orc_lambda_objective(5, orc._dp(lam), orc._dp(grad), orc._dp(nu), orc._dp(c), orc._dp(st),
                               orc._dp(mu), orc._dp(S), arith)
    np.testing.assert_allclose(v, g["expected_value"], rtol=RTOL)
    np.testing.assert_allclose(grad, g["expected_grad"], rtol=RTOL)


@pytest.mark.parametrize("arith", ARITHS)
def test_nu_objective(golden, arith):
    g = golden["nu_objective"]
    L = orc.lib()
    lam, nu, mu = (np.asarray(g[k], float) for k in ("lambda", "nu", "mu"))
    c = np.array([13 / g["zeta"][0]] * 2 + [7 / g["zeta"][1]] * 3)
    S = np.eye(5)
    grad = np.zeros(5)
    v = L.orc_nu_objective(5, orc._dp(nu), orc._dp(grad), orc._dp(lam), orc._dp(c), orc._dp(mu),
                           orc._dp(S), arith)
    np.testing.assert_allclose(v, g["expected_value"], rtol=RTOL)
    np.testing.assert_allclose(grad, g["expected_grad"], rtol=RTOL)


@pytest.mark.parametrize("arith", ARITHS)
def test_update_zeta(golden, arith):
    g = golden["update_zeta"]
    m = _toy(golden, arith)
    m.lam[:] = np.asarray(g["lambda"], float); m.nu[:] = np.asarray(g["nu"], float)
    m.L.orc_mmctm_update_zeta(m.p, 0)
    np.testing.assert_allclose(m.zeta[0], g["expected_d1"], rtol=RTOL)


@pytest.mark.parametrize("arith", ARITHS)
def test_update_theta(golden, arith):
    g = golden["update_theta"]
    m = _toy(golden, arith)
    m.lam[:] = np.asarray(g["lambda"], float)
    m.gamma[:] = np.concatenate([np.ravel(g["gamma"][0]), np.ravel(g["gamma"][1])]).astype(float)
    m.L.orc_mmctm_update_Elnphi(m.p)
    m.L.orc_mmctm_update_theta(m.p, 0)
    np.testing.assert_allclose(m.theta(0)[0:2].T, g["expected_theta_d1_m1"], rtol=RTOL)
    np.testing.assert_allclose(m.theta(0)[0:2].sum(axis=1), 1.0, rtol=1e-15)
    m.L.orc_mmctm_update_theta(m.p, 1)
    np.testing.assert_allclose(m.theta(1)[2:4].T, g["expected_theta_d2_m2"], rtol=RTOL)


@pytest.mark.parametrize("arith", ARITHS)
def test_update_mu_Sigma(golden, arith):
    m = _toy(golden, arith)
    g = golden["update_mu"]
    m.lam[:] = np.asarray(g["lambda"], float)
    m.L.orc_mmctm_update_mu(m.p)
    np.testing.assert_allclose(m.mu, g["expected"], rtol=RTOL)
    g = golden["update_Sigma"]
    m.lam[:] = np.asarray(g["lambda"], float); m.nu[:] = np.asarray(g["nu"], float)
    m.mu[:] = np.asarray(g["mu"], float)
    m.L.orc_mmctm_update_Sigma(m.p)
    np.testing.assert_allclose(m.Sigma, g["expected_Sigma"], rtol=RTOL)
    np.testing.assert_allclose(m.invSigma, g["expected_invSigma"], rtol=1e-12, atol=1e-14)


def test_update_gamma_det_product_form_equals_definition(golden):
    """Pinned specification: γ_kv = fma(E_kv, Σ_d exp(λ_dk) n_dv / Z_dv, α) (the (D x K)ᵀ(D x V) form, on
    the E-step's intermediates) must equal update_γ! (src/MMCTM.jl:224-240) on the θ that E-step stored."""
    m = _toy(golden, orc.ARITH_DET)
    for d in range(m.D):
        m.L.orc_mmctm_update_theta(m.p, d)
    ref = []
    for mod in range(m.M):
        rp, term, cnt = m._keep[mod]
        th = m.theta(mod)
        K, V = th.shape[1], m.V[mod]
        g = np.full((K, V), m.alpha[mod])
        for w in range(th.shape[0]):
            g[:, term[w]] += th[w] * cnt[w]
        ref.append(g.ravel())
    m.L.orc_mmctm_update_gamma(m.p)
    np.testing.assert_allclose(m.gamma, np.concatenate(ref), rtol=1e-13)


def test_update_gamma_Elnphi(golden):
    g = golden["update_gamma"]
    m = _toy(golden, orc.ARITH_LITERAL)
    _set_theta(m, 0, 0, g["theta"]["d1m1"]); _set_theta(m, 0, 1, g["theta"]["d2m1"])
    _set_theta(m, 1, 0, g["theta"]["d1m2"]); _set_theta(m, 1, 1, g["theta"]["d2m2"])
    m.L.orc_mmctm_update_gamma(m.p)
    exp = np.concatenate([np.ravel(g["expected"][0]), np.ravel(g["expected"][1])])
    np.testing.assert_allclose(m.gamma, exp, rtol=RTOL)
    g = golden["update_Elnphi"]
    m.gamma[0:4] = np.asarray(g["gamma_m1_k1"], float)
    m.L.orc_mmctm_update_Elnphi(m.p)
    np.testing.assert_allclose(m.Elnphi[0], g["expected_first"], rtol=1e-14)


def test_alpha_objective(golden):
    g = golden["alpha_objective"]
    L = orc.lib()
    grad = C.c_double()
    v = L.orc_alpha_objective(g["alpha"], C.byref(grad), g["sum_Elnphi"], g["K"], g["V"])
    np.testing.assert_allclose(v, g["expected_value"], rtol=RTOL)
    np.testing.assert_allclose(grad.value, g["expected_grad"], rtol=RTOL)


@pytest.mark.parametrize("arith", ARITHS)
def test_loglikelihoods(golden, arith):
    g = golden["loglikelihoods"]
    m = _toy(golden, arith)
    eta = np.asarray(g["eta"])
    m.lam[:, 0:2] = eta
    m.L.orc_mmctm_update_props(m.p)
    gam = np.asarray(g["gamma_m1"], float)
    m.phi[0:8] = (gam / gam.sum(axis=1, keepdims=True)).ravel()
    m.phi[8:] = 0.25
    ll = m.loglikelihoods()
    np.testing.assert_allclose(ll[0], g["expected_m1"], rtol=RTOL)


@pytest.mark.parametrize("arith", ARITHS)
def test_smoke_tests_of_reference(golden, arith):
    """test/mmctm.jl:92-101 (lambda changed, no NaN), :150-155 (nu > 0), :337-347 (elbo <= 0,
    fit returns one LL vector of length M)."""
    rng = np.random.default_rng(3)
    m = _toy(golden, arith, rng.integers(1, 101, 20))
    lam0 = np.array([1., 2, 3, 4, 1])
    m.lam[0] = lam0
    m.L.orc_mmctm_update_lambda(m.p, 0)
    assert not np.allclose(m.lam[0], lam0) and not np.isnan(m.lam[0]).any()
    m.mu[:] = [1, 1, 2, 2, 1]; m.lam[0] = lam0; m.nu[0] = [1, 1, 1, 2, 1]; m.zeta[0] = [2, 1]
    m.L.orc_mmctm_update_nu(m.p, 0)
    assert np.all(m.nu[0] > 0)
    m2 = _toy(golden, arith, rng.integers(1, 101, 20))
    assert m2.elbo()[0] <= 0.0
    ll = m2.fit(maxiter=1)
    assert ll.shape == (1, 2)


def test_special_functions(golden):
    L = orc.lib()
    for x, v in golden["special"]["digamma"].items():
        for f in (L.orc_digamma, L.orc_digamma_det):
            assert abs(f(float(x)) - v) <= 4e-15 * max(1.0, abs(v)), x
    for x, v in golden["special"]["lgamma"].items():
        assert abs(L.orc_lgamma(float(x)) - v) <= 1e-14 * max(1.0, abs(v)), x


def test_det_exp_log_accuracy():
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    L = orc.lib()
    rng = np.random.default_rng(0)
    for f, ref, xs in ((L.orc_exp, mp.exp, np.concatenate([rng.uniform(-30, 30, 3000), rng.uniform(-700, 700, 1000)])),
                       (L.orc_log, mp.log, np.concatenate([np.exp(rng.uniform(-30, 30, 3000)), rng.uniform(0.5, 2, 1000)]))):
        worst = 0.0
        for x in xs:
            y = f(float(x))
            worst = max(worst, float(abs((mp.mpf(y) - ref(mp.mpf(float(x)))) / np.spacing(abs(y)))))
        assert worst < 1.0, worst
    assert L.orc_exp(0.0) == 1.0 and L.orc_log(1.0) == 0.0
    assert L.orc_exp(-800.0) == 0.0 and L.orc_exp(710.0) == np.inf
    assert L.orc_log(0.0) == -np.inf and np.isnan(L.orc_log(-1.0))
    assert L.orc_exp(-745.0) == 5e-324 and abs(L.orc_log(5e-324) - (-744.4400719213812)) < 1e-12


# ---------------------------------------------------------------- LDA, test/lda.jl
def _lda(golden, arith, lam0=None):
    t = golden["lda_toy"]
    nested = [[x] for x in t["X"]]
    csr = from_nested(nested, 1)[0]
    l0 = np.arange(1, 5, dtype=float) if lam0 is None else lam0
    return orc.OracleLDA(t["K"], t["alpha"], t["eta"], t["V"], csr, l0, arith=arith)


@pytest.mark.parametrize("arith", ARITHS)
def test_lda_updates(golden, arith):
    m = _lda(golden, arith)
    g = golden["lda_update_phi"]
    m.Elntheta[:] = np.asarray(g["Elntheta"]).T          # golden is [k][d]
    m.Elnbeta[:] = np.asarray(g["Elnbeta"]).T            # golden is [v][k]
    m.L.orc_lda_update_phi(m.p)
    np.testing.assert_allclose(m.phi[0:2].T, g["expected_phi_d1"], rtol=RTOL)

    g = golden["lda_update_gamma"]
    m.phi[0:2] = np.asarray(g["phi_d1"]).T
    m.L.orc_lda_update_gamma(m.p)
    np.testing.assert_allclose(m.gamma[0], g["expected_gamma_d1"], rtol=RTOL)
    np.testing.assert_allclose(m.Elntheta[0], g["expected_Elntheta_d1"], rtol=1e-13)

    g = golden["lda_update_lambda"]
    m.phi[0:2] = np.asarray(g["phi"][0]).T; m.phi[2:4] = np.asarray(g["phi"][1]).T
    m.L.orc_lda_update_lambda(m.p)
    np.testing.assert_allclose(m.lam.T, g["expected_lambda_vk"], rtol=RTOL)
    np.testing.assert_allclose(m.Elnbeta.T, g["expected_Elnbeta_vk"], rtol=1e-13)


@pytest.mark.parametrize("arith", ARITHS)
def test_lda_elbo_negative_and_fit(golden, arith):
    m = _lda(golden, arith, np.array([3., 50, 17, 99]))
    assert m.elbo()[0] < 0.0                       # test/lda.jl:105-118
    ll = m.fit(maxiter=3)
    assert ll.shape == (3,) and np.all(np.isfinite(ll))

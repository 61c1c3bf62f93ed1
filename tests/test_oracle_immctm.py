"""The oracle's IMMCTM restatement (reference src/IMMCTM.jl) against the known answers of the
reference's own tests (test/immctm.jl, transcribed in tests/golden/immctm_known_answers.json), in
the literal and the pinned arithmetic; plus consistency of the pinned specification with the literal."""
import json
import os

import numpy as np
import pytest

import orc
from mmsig.counts import from_nested
from conftest import ROOT

G = json.load(open(os.path.join(ROOT, "tests", "golden", "immctm_known_answers.json")))
ARITHS = [orc.ARITH_LITERAL, orc.ARITH_DET]
RTOL = 1e-12


def _toy(arith, gammaf0=None):
    feats = [np.asarray(f) - 1 for f in G["features"]]
    counts = from_nested([[np.asarray(xm) for xm in xd] for xd in G["X"]], 2)
    T = sum(k * 4 for k in G["K"])
    g0 = np.arange(1, T + 1, dtype=float) if gammaf0 is None else np.asarray(gammaf0, float)
    return orc.OracleIMMCTM(G["K"], G["alpha"], feats, counts, g0, arith=arith)


def _flat(g):
    return [x for m in g for k in m for i in k for x in i]


@pytest.mark.parametrize("arith", ARITHS)
def test_ctor(arith):
    m = _toy(arith)
    c = G["ctor"]
    assert m.N().tolist() == c["N"] and m.I == c["I"] and m.J == c["J"] and list(m.V) == c["V"]
    assert m.MK == 5 and np.all(m.nu == 1.0) and np.all(m.lam == 0.0)
    assert np.array_equal(m.Sigma, np.eye(5)) and np.array_equal(m.invSigma, np.eye(5))


@pytest.mark.parametrize("arith", ARITHS)
def test_update_theta(arith):
    g = G["update_theta"]
    m = _toy(arith, _flat(g["gamma"]))
    m.lam[:] = np.asarray(g["lambda"], float)
    m.L.orc_mmctm_update_theta(m.p, 0)
    m.L.orc_mmctm_update_theta(m.p, 1)
    np.testing.assert_allclose(m.theta(0)[0:2].T, g["expected_theta_d1_m1"], rtol=RTOL)
    np.testing.assert_allclose(m.theta(1)[2:4].T, g["expected_theta_d2_m2"], rtol=RTOL)
    assert np.allclose(m.theta(0).sum(axis=1), 1.0) and (m.theta(0) >= 0).all()


def test_update_gamma_and_Elnphi():
    g = G["update_gamma"]
    m = _toy(orc.ARITH_LITERAL)
    m.theta(0)[0:2] = np.asarray(g["theta_d1_m1"]).T
    m.theta(0)[2:4] = np.asarray(g["theta_d2_m1"]).T
    m.L.orc_immctm_update_gamma(m.p)
    np.testing.assert_allclose(m.table(m.gammaf, 0, 0, 0), g["expected_m1_k1_i1"], rtol=RTOL)
    np.testing.assert_allclose(m.table(m.gammaf, 0, 0, 1), g["expected_m1_k1_i2"], rtol=RTOL)
    e = G["update_Elnphi"]
    m.table(m.gammaf, 0, 0, 0)[:] = e["gamma_m1_k1_i1"]
    m.L.orc_immctm_update_Elnphi(m.p)
    np.testing.assert_allclose(m.table(m.Elnphif, 0, 0, 0)[0], e["expected_first"], rtol=1e-14)


@pytest.mark.parametrize("arith", ARITHS)
def test_loglikelihood(arith):
    g = G["loglikelihood"]
    m = _toy(arith)
    for k in range(2):
        for i in range(2):
            m.table(m.gammaf, 0, k, i)[:] = g["gamma_m1"][k][i]
    m.L.orc_immctm_update_Elnphi(m.p)
    m.lam[:, 0:2] = np.asarray(g["eta"], float)
    m.L.orc_mmctm_update_props(m.p)
    np.testing.assert_allclose(m.loglikelihoods()[0], g["expected_m1"], rtol=RTOL)


def test_det_update_gamma_equals_definition():
    """Pinned specification (composite tables, product-form statistics) against update_γ! of
    src/IMMCTM.jl:197-221 on the θ the same E-step stored."""
    m = _toy(orc.ARITH_DET)
    for d in range(m.D):
        m.L.orc_mmctm_update_theta(m.p, d)
    ref = np.concatenate([np.full(int(m.K[mm]) * sum(m.J[mm]), G["alpha"][mm]) for mm in range(2)])
    off = 0
    for mod in range(2):
        rp, term, cnt = m._keep[mod]
        th = m.theta(mod)
        f = m._feats[mod]
        for w in range(th.shape[0]):
            for k in range(th.shape[1]):
                o = off + k * sum(m.J[mod])
                for i in range(f.shape[1]):
                    ref[o + f[term[w], i]] += th[w, k] * cnt[w]
                    o += m.J[mod][i]
        off += int(m.K[mod]) * sum(m.J[mod])
    m.L.orc_immctm_update_gamma(m.p)
    np.testing.assert_allclose(m.gammaf, ref, rtol=1e-13)


@pytest.mark.parametrize("arith", ARITHS)
def test_fit_runs_and_improves(arith):
    """test/immctm.jl "fit" (:343-348): one iteration returns one LL vector of length M; a few more
    iterations do not decrease the ELBO on a slightly larger random corpus."""
    m = _toy(arith)
    assert m.fit(maxiter=1).shape == (1, 2)
    rng = np.random.default_rng(0)
    feats = [np.stack(np.meshgrid(range(3), range(2), indexing="ij"), -1).reshape(-1, 2), np.arange(5).reshape(5, 1)]
    X = [[np.column_stack([np.arange(1, f.shape[0] + 1), rng.integers(1, 30, f.shape[0])]) for f in feats] for _ in range(40)]
    T = 2 * (3 + 2) + 2 * 5
    o = orc.OracleIMMCTM([2, 2], [0.1, 0.1], feats, from_nested(X, 2), rng.integers(1, 101, T).astype(float), arith=arith)
    o.fit(maxiter=3)
    e0 = o.elbo()[0]
    o.fit(maxiter=5)
    e1 = o.elbo()[0]
    assert np.isfinite(e0) and e1 > e0 - 1e-6 * abs(e0)
    # a one-feature IMMCTM is the MMCTM
    f1 = [np.arange(4).reshape(4, 1), np.arange(4).reshape(4, 1)]
    counts = from_nested([[np.asarray(xm) for xm in xd] for xd in G["X"]], 2)
    g0 = rng.integers(1, 101, 2 * 4 + 3 * 4).astype(float)
    a = orc.OracleIMMCTM(G["K"], G["alpha"], f1, counts, g0, arith=arith)
    b = orc.OracleMMCTM(G["K"], G["alpha"], [4, 4], counts, g0, arith=arith)
    ha, hb = a.fit(maxiter=4), b.fit(maxiter=4)
    np.testing.assert_allclose(ha, hb, rtol=1e-12)
    np.testing.assert_allclose(a.elbo()[0], b.elbo()[0], rtol=1e-12)

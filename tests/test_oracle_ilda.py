"""The oracle's ILDA restatement (reference src/ILDA.jl) against the known answers of the
reference's own tests (test/ilda.jl, transcribed in tests/golden/ilda_known_answers.json)."""
import json
import os

import numpy as np
import pytest

import orc
from mmsig.counts import from_nested
from conftest import ROOT

G = json.load(open(os.path.join(ROOT, "tests", "golden", "ilda_known_answers.json")))
ARITHS = [orc.ARITH_LITERAL, orc.ARITH_DET]


def _toy(arith, eta=None, lam0=None):
    feat = np.asarray(G["features"]) - 1
    X = [[np.asarray(x)] for x in G["X"]]
    csr = from_nested(X, 1)[0]
    l0 = np.arange(1, 2 * 4 + 1, dtype=float) if lam0 is None else lam0
    return orc.OracleILDA(G["K"], G["alpha"], G["eta"] if eta is None else eta, feat, csr, l0, arith=arith)


def _set_table(m, flat, tabs):
    """tabs[i] is the reference's J_i x K matrix of feature i."""
    for i, t in enumerate(tabs):
        t = np.asarray(t, float)
        for k in range(m.K):
            m.table(flat, k, i)[:] = t[:, k]


@pytest.mark.parametrize("arith", ARITHS)
def test_ctor_and_update_phi(arith):
    m = _toy(arith)
    assert m.I == G["ctor"]["I"] and m.J == G["ctor"]["J"] and np.all(m.gamma == 1.0)
    g = G["update_phi"]
    m.Elntheta[:] = np.asarray(g["Elntheta"]).T
    _set_table(m, m.Elnbetaf, g["Elnbeta"])
    m.L.orc_ilda_compose(m.p)
    m.L.orc_lda_update_phi(m.p)
    np.testing.assert_allclose(m.phi[0:2].T, g["expected_d1"], rtol=1e-12)
    np.testing.assert_allclose(m.phi[2:4].T, g["expected_d2"], rtol=1e-12)


@pytest.mark.parametrize("arith", ARITHS)
def test_update_gamma(arith):
    g = G["update_gamma"]
    m = _toy(arith)
    m.phi[0:2] = np.asarray(g["phi_d1"]).T
    m.L.orc_lda_update_gamma(m.p)
    np.testing.assert_allclose(m.gamma[0], g["expected_gamma_d1"], rtol=1e-12)
    np.testing.assert_allclose(m.Elntheta[0], g["expected_Elntheta_d1"], rtol=1e-12)


@pytest.mark.parametrize("arith", ARITHS)
def test_update_lambda(arith):
    g = G["update_lambda"]
    m = _toy(arith, eta=g["eta"])
    m.phi[0:2] = np.asarray(g["phi"][0]).T
    m.phi[2:4] = np.asarray(g["phi"][1]).T
    m.L.orc_lda_update_lambda(m.p)
    for i in range(2):
        for k in range(2):
            np.testing.assert_allclose(m.table(m.lambdaf, k, i), np.asarray(g["expected_lambda"][i])[:, k], rtol=1e-12)
            np.testing.assert_allclose(m.table(m.Elnbetaf, k, i), np.asarray(g["expected_Elnbeta"][i])[:, k], rtol=1e-12)


@pytest.mark.parametrize("arith", ARITHS)
def test_fit_and_one_feature_model_is_lda(arith):
    rng = np.random.default_rng(1)
    feat = np.stack(np.meshgrid(range(3), range(4), indexing="ij"), -1).reshape(-1, 2)
    X = [[np.column_stack([np.arange(1, 13), rng.integers(1, 20, 12)])] for _ in range(30)]
    csr = from_nested(X, 1)[0]
    m = orc.OracleILDA(3, 0.1, 0.1, feat, csr, rng.integers(1, 101, 3 * 7).astype(float), arith=arith)
    h = m.fit(maxiter=15)
    assert len(h) >= 11 and np.all(np.isfinite(h)) and h[-1] >= h[0] and np.isfinite(m.elbo()[0])
    # one feature whose values are the terms: the LDA itself
    f1 = np.arange(12).reshape(12, 1)
    l0 = rng.integers(1, 101, 3 * 12).astype(float)
    a = orc.OracleILDA(3, 0.1, 0.1, f1, csr, l0, arith=arith)
    b = orc.OracleLDA(3, 0.1, 0.1, 12, csr, l0, arith=arith)
    np.testing.assert_allclose(a.fit(maxiter=5), b.fit(maxiter=5), rtol=1e-12)
    np.testing.assert_allclose(a.elbo()[0], b.elbo()[0], rtol=1e-12)     # with one feature the `=` of :177 is harmless

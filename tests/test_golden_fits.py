"""Committed golden results of whole fits (tests/golden/config1_det.json): the oracle must still
reproduce them (CPU), and the CUDA path must reproduce them without consulting the oracle (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

import mmsig
from conftest import ROOT

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "config1_det.json")))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def _data(name, brca):
    g = GOLD[name]
    if name == "config1_brca_eu":
        return g, brca
    return g, mmsig.synth.generate(2000, g["K"], g["V"], key=mmsig.synth.DATA_KEY)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_golden_fit(name, brca):
    import orc
    g, counts = _data(name, brca)
    o = orc.OracleMMCTM(g["K"], [0.1] * len(g["K"]), g["V"], counts, mmsig.synth.init_gamma(g["K"], g["V"]),
                        arith=orc.ARITH_DET, nthreads=os.cpu_count() or 1)
    iters = 25 if name == "config1_brca_eu" else 6
    hist = o.fit(maxiter=iters, tol=1e-5)
    assert [[float(x).hex() for x in r] for r in hist] == g["ll_history_hex"]
    for k, h in g["sha256"].items():
        assert digest(getattr(o, k)) == h, k
    assert abs(o.elbo()[0] - g["elbo"]) <= 1e-12 * abs(g["elbo"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_reproduces_golden_fit(name, brca):
    g, counts = _data(name, brca)
    m = mmsig.MMCTM(g["K"], [0.1] * len(g["K"]), counts, V=g["V"], gamma0=mmsig.synth.init_gamma(g["K"], g["V"]))
    iters = 25 if name == "config1_brca_eu" else 6
    hist = m.fit(maxiter=iters, tol=1e-5, verbose=False)
    assert [[float(x).hex() for x in r] for r in hist] == g["ll_history_hex"]
    s = m.state()
    for k, h in g["sha256"].items():
        assert digest(s[k]) == h, k
    assert abs(m.elbo - g["elbo"]) <= 1e-12 * abs(g["elbo"])
    np.testing.assert_allclose(m.calculate_elbo()[1], g["elbo_terms"], rtol=1e-12)
    nn, nl = m.evals()
    assert int(nn.sum()) == g["evals_last_iteration"]["nu_sum"] and int(nl.sum()) == g["evals_last_iteration"]["lambda_sum"]
    m.close()


def test_pinned_arithmetic_stays_within_the_reference_own_sensitivity(brca):
    """the pinned-arithmetic oracle and the literal one (the reference's operation order,
    glibc exp/log) are the same algorithm up to roundings.  They agree to rounding level while no MMA stop
    decision has flipped, and afterwards drift apart no further than the reference drifts from itself
    under a 1e-15 perturbation of its start (ϕ: 4e-4, log-likelihood: 1e-4 relative after 20 iterations)."""
    import orc
    K, V = [7, 7], [96, 48]
    g0 = mmsig.synth.init_gamma(K, V)
    a, b = (orc.OracleMMCTM(K, [0.1, 0.1], V, brca, g0, arith=ar, nthreads=os.cpu_count() or 1)
            for ar in (orc.ARITH_LITERAL, orc.ARITH_DET))
    for it in range(12):
        la, lb = np.array(a.iterate()), np.array(b.iterate())
        rel = float(np.max(np.abs((la - lb) / la)))
        if it == 0:
            assert rel <= 1e-11 and np.max(np.abs(a.lam - b.lam)) <= 1e-5 and np.max(np.abs(a.phi - b.phi)) <= 1e-13
        assert rel <= 1e-4, (it, rel)
    assert np.max(np.abs(a.phi - b.phi)) <= 2e-3
    assert abs(a.elbo()[0] - b.elbo()[0]) <= 1e-4 * abs(b.elbo()[0])

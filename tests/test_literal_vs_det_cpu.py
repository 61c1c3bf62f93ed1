"""How far the pinned arithmetic (ORC_ARITH_DET: what the device computes bit for bit) is from the LITERAL
restatement of the reference (glibc exp / log, sequential sums, `/` and sqrt as written), one E+M iteration
from an IDENTICAL state.  These are the figures DESIGN.md section 2 and INTEGRATION.md print wherever
"bit-exact" is claimed; tests/test_gpu_mmctm.py asserts the same thresholds with the device in DET's place.

north_star asks for 1e-12 on phi, lambda, nu, mu, Sigma after one iteration.  That holds for phi and the
log-likelihood.  It cannot hold for lambda, nu (and mu, Sigma, which average them): they are where NLopt's
LD_MMA stops at xtol = 1e-4, and the stop is decided by the last bits of the objective; two arithmetics
that differ by roundings (two libm builds under the reference itself) stop 1e-11 (median) to 1e-8 apart
on the first iteration even when they take the same number of evaluations, and a few samples per thousand
take a different trace altogether."""
import numpy as np
import pytest

import orc
import mmsig
from util import small_synth, rel_err, norm_err

FIRST = dict(same=0.99, dlam=1e-7, dnu=1e-7, phi=1e-13, ll=1e-11, mu=1e-8, sigma=1e-7)
LATER = dict(same=0.95, phi=1e-13, ll=1e-6)


def _stats(a, b, la, lb):
    same = (a.nev_nu == b.nev_nu) & (a.nev_lambda == b.nev_lambda)
    dl = np.abs(a.lam - b.lam).max(axis=1)
    dn = (np.abs(a.nu - b.nu) / np.abs(a.nu)).max(axis=1)
    return dict(same=float(same.mean()), diverged=int((~same).sum()), dlam=float(dl[same].max()), dnu=float(dn[same].max()),
                dlam_median=float(np.median(dl[same])), dlam_diverged=float(dl[~same].max()) if (~same).any() else 0.0,
                phi=rel_err(b.phi, a.phi), ll=rel_err(lb, la), mu=norm_err(b.mu, a.mu), sigma=norm_err(b.Sigma, a.Sigma))


def _case(case, brca):
    if case == "brca":
        return [7, 7], [96, 48], brca
    K, V = [10, 8, 6], [96, 32, 83]
    return K, V, small_synth(2000, K, V, empty_frac=0.05)


@pytest.mark.parametrize("case", ["brca", "config4"])
def test_first_iteration(brca, case):
    K, V, counts = _case(case, brca)
    g0 = mmsig.synth.init_gamma(K, V)
    a = orc.OracleMMCTM(K, [0.1] * len(K), V, counts, g0, arith=orc.ARITH_LITERAL, nthreads=8)
    b = orc.OracleMMCTM(K, [0.1] * len(K), V, counts, g0, arith=orc.ARITH_DET, nthreads=8)
    st = _stats(a, b, a.iterate(), b.iterate())
    print("literal-vs-pinned, first iteration,", case, st)
    assert st["same"] >= FIRST["same"], st
    for k in ("dlam", "dnu", "phi", "ll", "mu", "sigma"):
        assert st[k] <= FIRST[k], (k, st)
    # and north_star's 1e-12 is NOT met on lambda even among same-trace samples: keep the claim honest
    assert st["dlam"] > 1e-12


@pytest.mark.parametrize("case,iters", [("brca", 12), ("config4", 6)])
def test_later_iteration_from_identical_state(brca, case, iters):
    K, V, counts = _case(case, brca)
    g0 = mmsig.synth.init_gamma(K, V)
    a = orc.OracleMMCTM(K, [0.1] * len(K), V, counts, g0, arith=orc.ARITH_LITERAL, nthreads=8)
    b = orc.OracleMMCTM(K, [0.1] * len(K), V, counts, g0, arith=orc.ARITH_DET, nthreads=8)
    for _ in range(iters):
        a.iterate()
    b.set_state(a.gamma.copy(), a.lam.copy(), a.nu.copy(), a.mu.copy(), a.Sigma.copy(), a.invSigma.copy())
    st = _stats(a, b, a.iterate(), b.iterate())
    print("literal-vs-pinned, iteration %d," % (iters + 1), case, st)
    assert st["same"] >= LATER["same"], st
    assert st["phi"] <= LATER["phi"] and st["ll"] <= LATER["ll"], st

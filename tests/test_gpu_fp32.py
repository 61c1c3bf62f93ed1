"""The optional FP32 mode (mmsig_config.precision = MMSIG_PRECISION_FP32, csrc/tile_f32.cuh): the tile passes in float,
sums over samples / the LD_MMA solves / the M-step tables in double.  north_star: "FP32 mode states its own looser
tolerance" -- stated here, against the oracle (LITERAL arithmetic for the LDA, which has no data-dependent branches;
the pinned DET arithmetic for the MMCTM) from the same state:

    one iteration   topic tables (ϕ, β, γ/λ statistics) 1e-5 relative, log-likelihood 1e-6 relative;
                    MMCTM: sumθ (the solver's input) 1e-5, and >= 90 % of the samples still take the same LD_MMA trace
    a 20-iteration fit   LDA: log-likelihood 1e-6, β 1e-6 absolute (no data-dependent branches: the error stays at float
                    rounding); MMCTM: log-likelihood and ELBO 1e-3 relative, ϕ 2e-2 absolute -- LD_MMA's stop decisions
                    amplify any perturbation (the reference's own sensitivity to a 1e-15 relative perturbation of γ₀ is
                    1e-6 on the ELBO and 4e-4 on ϕ after 20 iterations; float rounding is 1e-7).
Measured (B200): LDA 2e-7 / 5e-8 after one iteration, 7e-8 / 9e-8 after 20; MMCTM one iteration ϕ 2.4e-7, LL 6.7e-8, every
sample on the same LD_MMA trace, |Δλ| 7e-7; 20 iterations on brca-eu LL 2.1e-4, ϕ 3e-3.
"""
import numpy as np
import pytest

import orc
import mmsig
from util import small_synth, rel_err, oracle_mmctm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,V,D", [(20, 96, 3000), (7, 48, 500), (4, 5, 70)])
def test_lda_fp32_one_iteration_and_fit(K, V, D):
    csr = small_synth(D, [K], [V], empty_frac=0.05)[0]
    lam0 = mmsig.synth.init_lda_lambda(K, V)
    o = orc.OracleLDA(K, 0.1, 0.1, V, csr, lam0, arith=orc.ARITH_LITERAL, nthreads=8)
    g = mmsig.LDA(K, 0.1, 0.1, csr, V=V, lambda0=lam0, precision="fp32")
    ll_o, ll_g = o.iterate(), g.iterate()
    s = g.state()
    st = dict(lam=rel_err(s["lam"], o.lam), gamma=rel_err(s["gamma"], o.gamma), beta=rel_err(s["beta"], o.beta),
              ll=abs(ll_g - ll_o) / abs(ll_o))
    print("LDA fp32, one iteration, K=%d V=%d D=%d:" % (K, V, D), st)
    assert st["lam"] <= 1e-5 and st["gamma"] <= 1e-5 and st["beta"] <= 1e-5 and st["ll"] <= 1e-6, st
    for _ in range(19):
        ll_o, ll_g = o.iterate(), g.iterate()
    s = g.state()
    st = dict(beta_abs=float(np.max(np.abs(s["beta"] - o.beta))), ll=abs(ll_g - ll_o) / abs(ll_o))
    print("LDA fp32, 20 iterations:", st)
    assert st["ll"] <= 1e-6 and st["beta_abs"] <= 1e-6, st
    g.close()


@pytest.mark.parametrize("case", ["brca", "config4"])
def test_mmctm_fp32_one_iteration_and_fit(brca, case):
    if case == "brca":
        K, V, counts = [7, 7], [96, 48], brca
    else:
        K, V = [10, 8, 6], [96, 32, 83]
        counts = small_synth(2000, K, V, empty_frac=0.05)
    alpha = [0.1] * len(K)
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, alpha, V, counts, g0)
    g = mmsig.MMCTM(K, alpha, counts, V=V, gamma0=g0, precision="fp32")
    ll_o, ll_g = o.iterate(), g.iterate()
    s = g.state()
    nn, nl = g.evals()
    same = float(((nn == o.nev_nu) & (nl == o.nev_lambda)).mean())
    st = dict(phi=rel_err(s["phi"], o.phi), gamma=rel_err(s["gamma"], o.gamma), ll=rel_err(ll_g, ll_o), same_trace=same,
              dlam=float(np.max(np.abs(s["lam"] - o.lam))), mu=float(np.max(np.abs(s["mu"] - o.mu))))
    print("MMCTM fp32, one iteration,", case, st)
    assert st["phi"] <= 1e-5 and st["gamma"] <= 1e-5 and st["ll"] <= 1e-6, st
    assert st["same_trace"] >= 0.90 and st["dlam"] <= 1e-3, st          # λ: within LD_MMA's own x-tolerance (1e-4) x 10
    hist_o = [ll_o] + [o.iterate() for _ in range(19)]
    hist_g = [ll_g] + [g.iterate() for _ in range(19)]
    s = g.state()
    st = dict(ll=rel_err(hist_g[-1], hist_o[-1]), phi_abs=float(np.max(np.abs(s["phi"] - o.phi))))
    print("MMCTM fp32, 20 iterations,", case, st)
    assert st["ll"] <= 1e-3 and st["phi_abs"] <= 2e-2, st
    eg, eo = g.calculate_elbo()[0], o.elbo()[0]
    print("   ELBO rel diff", abs(eg - eo) / abs(eo))
    assert abs(eg - eo) <= 1e-3 * abs(eo)
    g.close()


def test_fp32_mode_is_opt_in_and_leaves_fp64_bits_alone(brca):
    """The default handle is FP64 and bit-exact; a bad precision value is refused."""
    K, V = [7, 7], [96, 48]
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, [0.1, 0.1], V, brca, g0)
    g = mmsig.MMCTM(K, [0.1, 0.1], brca, V=V, gamma0=g0)
    assert np.array_equal(g.iterate(), o.iterate())
    g.close()
    with pytest.raises(mmsig.capi.MmsigError):
        mmsig.capi.Handle(precision=7)

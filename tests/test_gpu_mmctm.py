"""CUDA path (through the C ABI) against the oracle: MMCTM / CTM.

Tolerances (north_star): one E+M iteration from an identical state within 1e-12 relative on
ϕ, λ, ν, μ, Σ, ELBO; a fixed-iteration fit within 1e-8 relative ELBO.  Against the oracle's
pinned arithmetic (ORC_ARITH_DET) the per-sample results are required to be BIT-EXACT, because
anything looser lets MMA's branch decisions diverge."""
import os

import numpy as np
import pytest

import orc
import mmsig
from mmsig.counts import from_nested
from util import oracle_mmctm, small_synth, rel_err, norm_err

pytestmark = pytest.mark.gpu

TOL_ITER = 1e-12
TOL_FIT_ELBO = 1e-8


def _pair(K, alpha, V, counts, gamma0, stop_rule=0):
    o = oracle_mmctm(K, alpha, V, counts, gamma0, stop_rule=stop_rule)
    g = mmsig.MMCTM(K, alpha, counts, V=V, gamma0=gamma0, stop_rule=stop_rule)
    return o, g


def _check_iteration(o, g, ll_o, ll_g, exact=True):
    s = g.state()
    nn, nl = g.evals()
    assert np.array_equal(nn, o.nev_nu) and np.array_equal(nl, o.nev_lambda), "MMA evaluation counts differ"
    if exact:
        assert np.array_equal(s["lam"], o.lam), "lambda not bit-exact: max abs %.3e" % np.abs(s["lam"] - o.lam).max()
        assert np.array_equal(s["nu"], o.nu), "nu not bit-exact"
        assert np.array_equal(s["zeta"], o.zeta), "zeta not bit-exact"
        assert np.array_equal(s["gamma"], o.gamma), "gamma not bit-exact: rel %.3e" % rel_err(s["gamma"], o.gamma)
        assert np.array_equal(s["mu"], o.mu), "mu not bit-exact"
        assert np.array_equal(s["Sigma"], o.Sigma), "Sigma not bit-exact"
        assert np.array_equal(s["invSigma"], o.invSigma), "invSigma not bit-exact"
        assert np.array_equal(s["Elnphi"], o.Elnphi) and np.array_equal(s["phi"], o.phi)
        assert np.array_equal(s["props"], o.props)
        assert np.array_equal(ll_g, ll_o), "log-likelihoods not bit-exact"
    assert rel_err(s["lam"], o.lam) <= TOL_ITER or norm_err(s["lam"], o.lam) <= TOL_ITER
    assert rel_err(s["nu"], o.nu) <= TOL_ITER
    assert rel_err(s["phi"], o.phi) <= TOL_ITER
    assert norm_err(s["mu"], o.mu) <= TOL_ITER
    assert norm_err(s["Sigma"], o.Sigma) <= TOL_ITER
    assert rel_err(ll_g, ll_o) <= TOL_ITER


def test_toy_corpus_one_iteration(golden):
    t = golden["mmctm_toy"]
    counts = from_nested(t["X"], 2)
    g0 = np.random.default_rng(1).integers(1, 101, 20).astype(float)
    o, g = _pair(t["K"], t["alpha"], [4, 4], counts, g0)
    # constructor state: Elnphi, zeta (src/MMCTM.jl:78-86)
    s = g.state(props=False)
    assert np.array_equal(s["Elnphi"], o.Elnphi) and np.array_equal(s["zeta"], o.zeta)
    ll_o, ll_g = o.iterate(), g.iterate()
    _check_iteration(o, g, ll_o, ll_g)
    eo, to = o.elbo()
    eg, tg = g.calculate_elbo()
    assert rel_err(tg, to) <= TOL_ITER and abs(eg - eo) <= TOL_ITER * abs(eo)
    g.close()


def test_brca_one_iteration_and_elbo(brca):
    K, alpha, V = [7, 7], [0.1, 0.1], [96, 48]
    g0 = mmsig.synth.init_gamma(K, V)
    o, g = _pair(K, alpha, V, brca, g0)
    ll_o, ll_g = o.iterate(), g.iterate()
    _check_iteration(o, g, ll_o, ll_g)
    eo, to = o.elbo()
    eg, tg = g.calculate_elbo()
    assert rel_err(tg, to) <= TOL_ITER, (tg, to)
    assert abs(eg - eo) <= TOL_ITER * abs(eo)
    for m in range(2):
        np.testing.assert_allclose(g.theta(m), o.theta(m), rtol=1e-13)
    g.close()


def test_brca_fit_config1(brca):
    """config 1: MMCTM([7,7],[0.1,0.1]) on brca-eu, fit!(tol=1e-5) (README.md:18-26)."""
    K, alpha, V = [7, 7], [0.1, 0.1], [96, 48]
    g0 = mmsig.synth.init_gamma(K, V)
    o, g = _pair(K, alpha, V, brca, g0)
    ho = o.fit(maxiter=30, tol=1e-5)
    hg = g.fit(maxiter=30, tol=1e-5, verbose=False)
    assert hg.shape == ho.shape and g.converged == o.converged
    assert np.array_equal(hg, ho), "LL history differs: %.3e" % rel_err(hg, ho)
    _check_iteration(o, g, ho[-1], hg[-1])
    eo, _ = o.elbo()
    assert abs(g.elbo - eo) <= TOL_FIT_ELBO * abs(eo)
    assert abs(g.elbo - eo) <= TOL_ITER * abs(eo)
    assert np.array_equal(g.ll, ho[-1])
    g.close()


@pytest.mark.parametrize("device_rule", ["1", "0"])
def test_fit_stops_exactly_where_the_reference_rule_fires(brca, monkeypatch, device_rule):
    """Beyond iteration 10 the iterations of fit are enqueued in batches of 8 and the stopping rule (src/MMCTM.jl:485) is
    evaluated on the device; when it fires in iteration j the later kernels of the batch return at once.  For tolerances
    that make the rule fire first at different positions of a batch (odd and even numbers of skipped iterations: the
    host-side lambda buffer swap has to be undone for an odd count) the history, the iteration count and the state are
    the oracle's, bit for bit -- the state of iteration j, not of the last enqueued one."""
    monkeypatch.setenv("MMSIG_DEVICE_RULE", device_rule)
    K, alpha, V = [7, 7], [0.1, 0.1], [96, 48]
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, alpha, V, brca, g0)
    full = o.fit(maxiter=30, tol=0.0)
    r = np.max(np.abs(full[:-1] - full[1:]) / np.abs(full[1:]), axis=1)        # r[i]: change entering iteration i + 2
    tols = []
    for j in (11, 12, 14, 17, 18, 19, 22):                                     # positions 0, 1, 3, 6, 7 of the first batch; 0, 3 of the second
        rj, prev = r[j - 2], r[9:j - 2]
        if prev.size == 0 or rj < prev.min():                                  # the change entering iteration j is a new minimum
            tols.append((j, float(rj * 1.0000001 if prev.size == 0 else 0.5 * (rj + prev.min()))))
    assert len(tols) >= 5, tols
    for j, tol in tols:
        oo = oracle_mmctm(K, alpha, V, brca, g0)
        ho = oo.fit(maxiter=30, tol=tol)
        assert len(ho) == j and oo.converged, (j, tol, len(ho))
        g = mmsig.MMCTM(K, alpha, brca, V=V, gamma0=g0)
        hg = g.fit(maxiter=30, tol=tol, verbose=False)
        assert hg.shape == ho.shape and g.converged, (j, hg.shape, ho.shape)
        assert np.array_equal(hg, ho)
        _check_iteration(oo, g, ho[-1], hg[-1])
        eo = oo.elbo()[0]                      # reads the sumθ, ζ and λ_prev of iteration j (a θ pass enqueued after it must not touch them)
        assert abs(g.elbo - eo) <= TOL_ITER * abs(eo), (j, g.elbo, eo)
        ll_next_o, ll_next_g = oo.iterate(), g.iterate()                       # and the handle goes on from there
        assert np.array_equal(ll_next_g, ll_next_o)
        g.close()


@pytest.mark.parametrize("K,V,D,empty", [([10, 8, 6], [96, 32, 83], 1500, 0.05),     # config 4 shape
                                         ([10], [96], 1200, 0.0),                    # config 3: CTM
                                         ([7, 7], [96, 32], 800, 0.1),               # config 5 shape
                                         ([1, 2], [5, 3], 64, 0.3),                  # degenerate sizes
                                         ([16, 16], [40, 7], 300, 0.0),              # MK = 32
                                         ([12, 12, 10], [96, 32, 83], 300, 0.05),    # MK = 34: two coordinates per lane
                                         ([20, 20], [30, 10], 200, 0.0),             # MK = 40
                                         ([32, 32], [40, 33], 150, 0.0)])            # MK = 64
def test_synthetic_three_iterations(K, V, D, empty):
    counts = small_synth(D, K, V, empty_frac=empty)
    alpha = [0.1] * len(K)
    g0 = mmsig.synth.init_gamma(K, V)
    o, g = _pair(K, alpha, V, counts, g0)
    for it in range(3):
        ll_o, ll_g = o.iterate(), g.iterate()
        _check_iteration(o, g, ll_o, ll_g)
    eo, to = o.elbo()
    eg, tg = g.calculate_elbo()
    assert abs(eg - eo) <= TOL_ITER * abs(eo), (tg, to)
    g.close()


def test_stop_rule_nlopt26_and_no_sigma_update(brca):
    K, alpha, V = [7, 7], [0.1, 0.1], [96, 48]
    g0 = mmsig.synth.init_gamma(K, V)
    o, g = _pair(K, alpha, V, brca, g0, stop_rule=1)
    for it in range(2):
        ll_o, ll_g = o.iterate(updateSigma=(it == 0)), g.iterate(updateSigma=(it == 0))
        _check_iteration(o, g, ll_o, ll_g)
    g.close()


def test_set_state_continues_a_fit(brca):
    """fit! is re-entrant: state downloaded from one handle and uploaded into another continues
    identically (src/MMCTM.jl: all state lives in the struct)."""
    K, alpha, V = [7, 7], [0.1, 0.1], [96, 48]
    g0 = mmsig.synth.init_gamma(K, V)
    a = mmsig.MMCTM(K, alpha, brca, V=V, gamma0=g0)
    a.iterate(); a.iterate()
    s = a.state()
    b = mmsig.MMCTM(K, alpha, brca, V=V, gamma0=s["gamma"])
    b.set_state(s["gamma"], lam=s["lam"], nu=s["nu"], mu=s["mu"], Sigma=s["Sigma"], invSigma=s["invSigma"])
    la, lb = a.iterate(), b.iterate()
    assert np.array_equal(la, lb)
    sa, sb = a.state(), b.state()
    for k in ("lam", "nu", "gamma", "mu", "Sigma"):
        assert np.array_equal(sa[k], sb[k]), k
    a.close(); b.close()


# Measured distance between the pinned arithmetic (what the device computes, bit for bit) and the LITERAL
# restatement of the reference (glibc exp / log, sequential sums), one E+M iteration from an identical state
# (DESIGN.md section 2 holds the table; tests/test_literal_vs_det_cpu.py asserts the same figures on the CPU).
# First iteration: brca-eu 5 of 560 samples take another MMA trace, config-4 shape none of 2000.
LITERAL_FIRST = dict(same=0.99, dlam=1e-7, dnu=1e-7, phi=1e-13, ll=1e-11, mu=1e-8, sigma=1e-7)
# From the literal oracle's state after 6 / 12 iterations: the conditioning of MMA's stop grows with the fit
# (cond(Sigma)), more samples flip; phi (a function of the OLD lambda only) and the LL stay tight.
LITERAL_LATER = dict(same=0.95, phi=1e-13, ll=1e-6)


def _literal_stats(o, g, ll_o, ll_g):
    nn, nl = g.evals()
    same = (nn == o.nev_nu) & (nl == o.nev_lambda)
    s = g.state()
    dl = np.abs(s["lam"] - o.lam).max(axis=1)
    dn = (np.abs(s["nu"] - o.nu) / np.abs(o.nu)).max(axis=1)
    return dict(same=float(same.mean()), diverged=int((~same).sum()), dlam=float(dl[same].max()), dnu=float(dn[same].max()),
                phi=rel_err(s["phi"], o.phi), ll=rel_err(ll_g, ll_o), mu=norm_err(s["mu"], o.mu),
                sigma=norm_err(s["Sigma"], o.Sigma))


def _assert_literal(st, lim):
    assert st["same"] >= lim["same"], st
    for k in lim:
        if k != "same":
            assert st[k] <= lim[k], (k, st)


@pytest.mark.parametrize("case", ["brca", "config4"])
def test_against_literal_oracle_statistics(brca, case):
    """Device path against the LITERAL oracle, first iteration from the constructor state, at the measured
    figures: >= 99 % of the samples take the same MMA trace (same evaluation counts); among those
    |dlambda| <= 1e-7, rel dnu <= 1e-7; phi <= 1e-13, LL <= 1e-11, mu <= 1e-8, Sigma <= 1e-7 (north_star's 1e-12
    holds for phi and the LL only: lambda, nu are where LD_MMA stops at xtol 1e-4, not an optimum)."""
    if case == "brca":
        K, V, counts = [7, 7], [96, 48], brca
    else:
        K, V = [10, 8, 6], [96, 32, 83]
        counts = small_synth(2000, K, V, empty_frac=0.05)
    alpha = [0.1] * len(K)
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, alpha, V, counts, g0, arith=orc.ARITH_LITERAL)
    g = mmsig.MMCTM(K, alpha, counts, V=V, gamma0=g0)
    ll_o, ll_g = o.iterate(), g.iterate()
    st = _literal_stats(o, g, ll_o, ll_g)
    print("literal-vs-device, first iteration,", case, st)
    _assert_literal(st, LITERAL_FIRST)
    g.close()


@pytest.mark.parametrize("case,iters", [("brca", 12), ("config4", 6)])
def test_against_literal_oracle_later_iteration(brca, case, iters):
    """The same comparison from the LITERAL oracle's own state after `iters` iterations (identical state on
    both sides through set_state): what a user of the reference would see between two builds of it."""
    if case == "brca":
        K, V, counts = [7, 7], [96, 48], brca
    else:
        K, V = [10, 8, 6], [96, 32, 83]
        counts = small_synth(2000, K, V, empty_frac=0.05)
    alpha = [0.1] * len(K)
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, alpha, V, counts, g0, arith=orc.ARITH_LITERAL)
    for _ in range(iters):
        o.iterate()
    g = mmsig.MMCTM(K, alpha, counts, V=V, gamma0=g0)
    g.set_state(o.gamma.copy(), lam=o.lam.copy(), nu=o.nu.copy(), mu=o.mu.copy(), Sigma=o.Sigma.copy(), invSigma=o.invSigma.copy())
    ll_o, ll_g = o.iterate(), g.iterate()
    st = _literal_stats(o, g, ll_o, ll_g)
    print("literal-vs-device, iteration %d," % (iters + 1), case, st)
    _assert_literal(st, LITERAL_LATER)
    g.close()


def test_bad_input_is_rejected():
    rp = np.array([0, 2, 3], np.int64)
    with pytest.raises(mmsig.capi.MmsigError):          # term out of range
        mmsig.MMCTM([2], [0.1], [(rp, np.array([0, 9, 1], np.int32), np.array([1, 1, 1], np.int32))], V=[4])
    with pytest.raises(mmsig.capi.MmsigError):          # unsorted row
        mmsig.MMCTM([2], [0.1], [(rp, np.array([3, 1, 1], np.int32), np.array([1, 1, 1], np.int32))], V=[4])
    with pytest.raises(mmsig.capi.MmsigError):          # zero count
        mmsig.MMCTM([2], [0.1], [(rp, np.array([0, 1, 1], np.int32), np.array([1, 0, 1], np.int32))], V=[4])
    with pytest.raises(mmsig.capi.MmsigError):          # sum(K) > 64
        mmsig.MMCTM([32, 32, 1], [0.1] * 3, [(rp, np.array([0, 1, 1], np.int32), np.array([1, 1, 1], np.int32))] * 3, V=[4, 4, 4])


def test_restarts_config5_shape():
    """config 5 (shape): independent restarts on resident counts, arg-max ELBO; the handle ends up
    holding the best restart's state."""
    K, V, D, R = [7, 7], [96, 32], 400, 4
    counts = small_synth(D, K, V)
    rng = np.random.default_rng(5)
    g0s = rng.integers(1, 101, size=(R, sum(k * v for k, v in zip(K, V)))).astype(float)
    g = mmsig.MMCTM(K, [0.1, 0.1], counts, V=V, gamma0=g0s[0])
    elbo, ll, nit, best = g.fit_restarts(g0s, maxiter=12, tol=1e-4)
    eo, states = [], []
    for r in range(R):
        o = oracle_mmctm(K, [0.1, 0.1], V, counts, g0s[r])
        h = o.fit(maxiter=12, tol=1e-4)
        eo.append(o.elbo()[0])
        states.append((o.lam.copy(), o.phi.copy(), h[-1].copy(), len(h)))
    eo = np.asarray(eo)
    assert rel_err(elbo, eo) <= TOL_ITER
    assert best == int(np.argmax(eo))
    assert np.array_equal(nit, [s[3] for s in states])
    s = g.state()
    assert np.array_equal(s["lam"], states[best][0]) and np.array_equal(s["phi"], states[best][1])
    assert np.array_equal(ll[best], states[best][2]) and np.array_equal(g.ll, states[best][2])
    assert abs(g.calculate_elbo()[0] - eo[best]) <= TOL_ITER * abs(eo[best])
    g.close()


def test_auto_alpha(brca):
    """fit!(autoα=true): update_α! (src/MMCTM.jl:252-269) after update_γ!, 1-D LD_MMA on the host."""
    K, alpha, V = [7, 7], [0.1, 0.1], [96, 48]
    g0 = mmsig.synth.init_gamma(K, V)
    o, g = _pair(K, alpha, V, brca, g0)
    ho = o.fit(maxiter=4, tol=1e-5, autoalpha=True)
    hg = g.fit(maxiter=4, tol=1e-5, verbose=False, autoalpha=True)
    assert np.array_equal(g.alpha, o.alpha) and not np.allclose(g.alpha, alpha)     # test/mmctm.jl:291
    assert np.array_equal(hg, ho)
    _check_iteration(o, g, ho[-1], hg[-1])
    eo, to = o.elbo()
    eg, tg = g.calculate_elbo()
    assert abs(eg - eo) <= TOL_ITER * abs(eo), (tg, to)
    g.close()


def test_two_stage_restart_orchestration():
    """scripts/run_mmctm.jl:163-182 on resident counts, against the same procedure on the oracle."""
    K, V, D, R = [4, 3], [30, 12], 300, 3
    counts = small_synth(D, K, V, seed=31)
    rng = np.random.default_rng(9)
    G = sum(k * v for k, v in zip(K, V))
    g0s = rng.integers(1, 101, size=(R, G)).astype(float)
    g = mmsig.MMCTM(K, [0.1, 0.1], counts, V=V, gamma0=g0s[0])
    out = mmsig.restarts.fit_model(g, g0s, maxiter=15)
    ll1, gam = [], []
    for r in range(R):
        o = oracle_mmctm(K, [0.1, 0.1], V, counts, g0s[r])
        h = o.fit(maxiter=15, tol=1e-4)
        ll1.append(h[-1].copy()); gam.append(o.gamma.copy())
    ll1 = np.asarray(ll1)
    assert np.array_equal(out["stage1_ll"], ll1)
    win = np.argmax(ll1, axis=0)
    assert out["winners"].tolist() == win.tolist()
    g2 = np.concatenate([gam[win[0]][:120], gam[win[1]][120:]])
    o2 = oracle_mmctm(K, [0.1, 0.1], V, counts, g2)
    h2 = o2.fit(maxiter=15, tol=1e-5)
    assert np.array_equal(out["stage2_ll"], h2[-1])
    assert np.array_equal(g.state()["lam"], o2.lam)
    g.close()


@pytest.mark.parametrize("variant", ["lean8", "lean4", "warp"])
def test_solver_layouts_are_bit_identical(monkeypatch, variant):
    """MMSIG_SOLVE selects the lane layout of the LD_MMA kernels for sum(K) <= 32: lean8 / lean4 (csrc/mmctm_lean.cuh:
    one kernel per phase, 8 / 4 lanes per sample, several coordinates per lane) or warp (k_solve / k_solve_pack: one
    kernel, one coordinate per lane).  Same arithmetic, so every layout is bit-identical to the oracle, per-sample
    evaluation counts included; edge shapes: sum(K) = 32 (no padding lane), 5 (mostly padding), empty rows."""
    monkeypatch.setenv("MMSIG_SOLVE", variant)
    for K, V, D in (([10, 8, 6], [96, 32, 83], 700), ([16, 16], [40, 7], 300), ([9, 8], [30, 20], 300), ([10], [96], 500),
                    ([7, 7], [96, 32], 400), ([3, 2], [10, 6], 67), ([12, 12, 5], [30, 20, 9], 200)):
        counts = small_synth(D, K, V, empty_frac=0.05)
        g0 = mmsig.synth.init_gamma(K, V)
        o, g = _pair(K, [0.1] * len(K), V, counts, g0)
        for _ in range(3):
            ll_o, ll_g = o.iterate(), g.iterate()
            _check_iteration(o, g, ll_o, ll_g)
        g.close()
    # stop rule of NLopt <= 2.6 and a starting nu below LD_MMA's lower bound (evaluated as given, then clamped)
    K, V = [10, 8, 6], [96, 32, 83]
    counts = small_synth(300, K, V, empty_frac=0.05)
    g0 = mmsig.synth.init_gamma(K, V)
    o, g = _pair(K, [0.1] * 3, V, counts, g0, stop_rule=1)
    nu0 = np.ones((300, 24)); nu0[::7, ::5] = 1e-9
    lam0 = np.zeros((300, 24))
    o.set_state(g0, lam0, nu0, np.zeros(24), np.eye(24), np.eye(24))
    g.set_state(g0, lam=lam0, nu=nu0)
    for _ in range(2):
        ll_o, ll_g = o.iterate(), g.iterate()
        _check_iteration(o, g, ll_o, ll_g)
    g.close()


@pytest.mark.parametrize("tiles", ["csr", "dense"])
def test_count_tile_formats_are_bit_identical(monkeypatch, tiles):
    """MMSIG_TILES selects what the tile kernels read: the CSR records (clear + scatter per tile) or the dense count
    tiles staged by cp.async.bulk + mbarrier (csrc/tile_stage.cuh; the default above 40 % density).  Same counts, same
    arithmetic: both are bit-identical to the oracle, on dense and on sparse data, with a ragged last tile (D not a
    multiple of 32), empty rows, one-warp (V <= 32) and multi-warp blocks, and through the chunked fit_host."""
    monkeypatch.setenv("MMSIG_TILES", tiles)
    for K, V, D, sparsify in (([10, 8, 6], [96, 32, 83], 715, 0.0), ([7, 7], [96, 48], 333, 0.8), ([3, 2], [10, 6], 67, 0.0),
                              ([12, 5], [130, 20], 97, 0.5)):
        counts = small_synth(D, K, V, empty_frac=0.05)
        if sparsify:                                   # drop most cells: the density rule alone would pick CSR here
            rng = np.random.default_rng(3)
            out = []
            for (rp, t, c), v in zip(counts, V):
                dense = np.zeros((D, v), dtype=np.int64)
                for d in range(D):
                    dense[d, t[rp[d]:rp[d + 1]]] = c[rp[d]:rp[d + 1]]
                dense[rng.random(dense.shape) < sparsify] = 0
                out.append(mmsig.counts.make_count_csr(dense.T))
            counts = out
        g0 = mmsig.synth.init_gamma(K, V)
        o, g = _pair(K, [0.1] * len(K), V, counts, g0)
        for _ in range(3):
            ll_o, ll_g = o.iterate(), g.iterate()
            _check_iteration(o, g, ll_o, ll_g)
        g.close()
    monkeypatch.setenv("MMSIG_PIPE_CHUNKS", "3")
    K, V, D = [10, 8, 6], [96, 32, 83], 1000
    counts = small_synth(D, K, V, empty_frac=0.05)
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, [0.1] * 3, V, counts, g0)
    llo = o.fit(maxiter=3, tol=1e-4)
    g = mmsig.MMCTM(K, [0.1] * 3, counts, V=V, gamma0=g0)
    hist, s = g.fit_host(counts, g0, maxiter=3)
    assert np.array_equal(hist, np.asarray(llo)) and np.array_equal(s["lam"], o.lam) and np.array_equal(s["phi"], o.phi)
    g.close()


@pytest.mark.parametrize("maxiter,chunks", [(1, 1), (1, 3), (4, 3), (14, 5)])
def test_fit_host_equals_the_four_calls(monkeypatch, maxiter, chunks):
    """mmsig_mmctm_fit_host (transfers pipelined behind the E-step, chunked launches accumulating
    their block partials) == set_data + set_state + fit + get_state, bit for bit, whether the loop
    ends by maxiter (outputs streamed out chunk by chunk) or by convergence."""
    K, V, D = [5, 4, 3], [96, 32, 83], 1500
    counts = small_synth(D, K, V, seed=3, empty_frac=0.05)
    g0 = mmsig.synth.init_gamma(K, V)
    a = mmsig.MMCTM(K, [0.1, 0.2, 0.3], counts, V=V, gamma0=g0)
    for _ in range(2):                       # a non-trivial starting state
        a.iterate()
    s0 = a.state()
    a.set_state(s0["gamma"], lam=s0["lam"], nu=s0["nu"], mu=s0["mu"], Sigma=s0["Sigma"], invSigma=s0["invSigma"])
    tol = 1e-3
    hist_a = a.fit(maxiter=maxiter, tol=tol, verbose=False)
    sa = a.state()
    monkeypatch.setenv("MMSIG_PIPE_CHUNKS", str(chunks))
    # a handle that held a different corpus before: fit_host must re-plan
    other = small_synth(700, K, V, seed=9)
    b = mmsig.MMCTM(K, [0.1, 0.2, 0.3], other, V=V, gamma0=g0)
    hist_b, sb = b.fit_host(counts, s0["gamma"], lam=s0["lam"], nu=s0["nu"], mu=s0["mu"], Sigma=s0["Sigma"],
                            invSigma=s0["invSigma"], maxiter=maxiter, tol=tol)
    assert np.array_equal(hist_a, hist_b) and a.converged == b.converged
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    assert b.calculate_elbo()[0] == a.elbo
    # and again on the now-matching shape (allocation reuse), from the constructor state
    c = mmsig.MMCTM(K, [0.1, 0.2, 0.3], counts, V=V, gamma0=g0)
    hist_c = c.fit(maxiter=maxiter, tol=tol, verbose=False)
    hist_d, sd = b.fit_host(counts, g0, maxiter=maxiter, tol=tol)
    sc = c.state()
    assert np.array_equal(hist_c, hist_d)
    for k in sc:
        assert np.array_equal(sc[k], sd[k]), k
    a.close(); b.close(); c.close()


def test_fit_host_packed_records_equal_unpacked():
    """mmsig_mmctm_fit_host_packed (4-byte records: term | count << 10) == mmsig_mmctm_fit_host, bit for bit; values that
    do not fit the record are refused by mmsig_pack_records."""
    K, V, D = [5, 4, 3], [96, 32, 83], 1500
    counts = small_synth(D, K, V, seed=3, empty_frac=0.05)
    g0 = mmsig.synth.init_gamma(K, V)
    a = mmsig.MMCTM(K, [0.1, 0.2, 0.3], counts, V=V, gamma0=g0)
    ha, sa = a.fit_host(counts, g0, maxiter=3)
    packed = [(r, mmsig.capi.pack_records(t, c)) for r, t, c in counts]
    hb, sb = a.fit_host(packed, g0, maxiter=3)
    assert np.array_equal(ha, hb)
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    grp = mmsig.MMCTMGroup(K, [0.1, 0.2, 0.3], counts, [0, 0], V=V, gamma0=g0)
    hc, sc = grp.fit_host(packed, g0, maxiter=3)
    assert np.array_equal(ha, hc)
    for k in sa:
        assert np.array_equal(sa[k], sc[k]), k
    with pytest.raises(mmsig.capi.MmsigError):
        mmsig.capi.pack_records(np.array([3, 1024], np.int32), np.array([1, 1], np.int32))
    with pytest.raises(mmsig.capi.MmsigError):
        mmsig.capi.pack_records(np.array([3, 4], np.int32), np.array([1, 1 << 22], np.int32))
    bad = [(r.copy(), x.copy()) for r, x in packed]
    bad[0][1][5] = np.uint32(200 | (3 << 10))          # term 200 >= V[0] = 96
    with pytest.raises(mmsig.capi.MmsigError):
        a.fit_host(bad, g0, maxiter=1)
    a.close(); grp.close()


def test_fit_host_rejects_bad_counts():
    K, V, D = [3, 2], [10, 6], 64
    counts = small_synth(D, K, V, seed=5)
    g0 = mmsig.synth.init_gamma(K, V)
    m = mmsig.MMCTM(K, [0.1, 0.1], counts, V=V, gamma0=g0)
    bad = [(r.copy(), t.copy(), c.copy()) for r, t, c in counts]
    bad[1][1][0] = 99                                  # term out of range
    with pytest.raises(mmsig.capi.MmsigError):
        m.fit_host(bad, g0, maxiter=1)
    bad2 = [(r.copy(), t.copy(), c.copy()) for r, t, c in counts]
    bad2[0][0][40] = bad2[0][0][41] + 1                # row pointers not monotone
    with pytest.raises(mmsig.capi.MmsigError):
        m.fit_host(bad2, g0, maxiter=1)
    hist, _ = m.fit_host(counts, g0, maxiter=2)        # the handle recovers
    assert np.isfinite(hist).all()
    # terms far outside the tile (negative, aliasing into the slot-tag bits of a record, huge): fit_host
    # launches a chunk's E-step before it reads the validation flags, so the packing kernel has to
    # neutralise such a record, not only flag it (an out-of-bounds shared-memory write would poison the context)
    for t_bad in (-1, 65535, 70000, 0x7fff0000, -(2 ** 31)):
        bad3 = [(r.copy(), t.copy(), c.copy()) for r, t, c in counts]
        w = int(bad3[0][0][7 + 1]) - 1                 # last nonzero of row 7 (keeps the row ascending unless negative)
        bad3[0][1][w] = t_bad
        with pytest.raises(mmsig.capi.MmsigError):
            m.fit_host(bad3, g0, maxiter=1)
        hist, _ = m.fit_host(counts, g0, maxiter=1)    # no sticky error: the same handle still works
        assert np.isfinite(hist).all()
    m.close()

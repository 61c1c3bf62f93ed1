"""Parity at the sizes the bench path actually exercises: D = 1e5 at config-4 shape (thousands of tiles, the
atomic sample queue across every block of a full grid, 64-bit offsets into the records, the default chunking
of mmsig_mmctm_fit_host) against the DET oracle, bit for bit; and at D = 1e6 the size-independent
properties (a fit split over two handles sums to the same statistics as one; the order of the samples does
not change the tables)."""
import os

import numpy as np
import pytest

import orc
import mmsig
from util import oracle_mmctm, rel_err

pytestmark = pytest.mark.gpu

K, V, ALPHA = [10, 8, 6], [96, 32, 83], [0.1, 0.1, 0.1]


def _same(s, o, ll_g, ll_o):
    for k, ref in (("lam", o.lam), ("nu", o.nu), ("zeta", o.zeta), ("gamma", o.gamma), ("mu", o.mu), ("Sigma", o.Sigma),
                   ("invSigma", o.invSigma), ("Elnphi", o.Elnphi), ("phi", o.phi), ("props", o.props)):
        assert np.array_equal(s[k], ref), "%s not bit-exact at D=1e5: %.3e" % (k, np.abs(s[k] - ref).max())
    assert np.array_equal(ll_g, ll_o)


def test_config4_shape_100k_two_iterations_bit_exact():
    D = 100_000
    counts = mmsig.synth.generate(D, K, V, key=20261018)
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, ALPHA, V, counts, g0, nthreads=os.cpu_count() or 8)
    g = mmsig.MMCTM(K, ALPHA, counts, V=V, gamma0=g0)
    for _ in range(2):
        ll_o, ll_g = o.iterate(), g.iterate()
        nn, nl = g.evals()
        assert np.array_equal(nn, o.nev_nu) and np.array_equal(nl, o.nev_lambda)
        _same(g.state(), o, ll_g, ll_o)
    eo, eg = o.elbo()[0], g.calculate_elbo()[0]
    assert abs(eg - eo) <= 1e-12 * abs(eo)
    g.close()


def test_fit_host_default_chunking_100k_bit_exact():
    """mmsig_mmctm_fit_host with the chunking it picks itself (4 chunks at D = 1e5), from host buffers, two
    iterations: identical to the oracle and to the resident path."""
    D = 100_000
    assert "MMSIG_PIPE_CHUNKS" not in os.environ
    counts = mmsig.synth.generate(D, K, V, key=7)
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, ALPHA, V, counts, g0, nthreads=os.cpu_count() or 8)
    ho = o.fit(maxiter=2, tol=1e-4)
    small = mmsig.synth.generate(64, K, V, key=3)
    g = mmsig.MMCTM(K, ALPHA, small, V=V, gamma0=g0)           # a handle planned for another corpus
    hg, s = g.fit_host(counts, g0, maxiter=2, tol=1e-4)
    assert np.array_equal(hg, ho)
    _same(s, o, hg[-1], ho[-1])
    g.close()


def test_million_samples_sharding_and_order_properties():
    """D = 1e6 (BASELINE config 4 size), no oracle: (1) the exactly-accumulated statistics of one handle over all
    samples equal those of the same samples cut into two handles and added (gamma - alpha is additive over
    samples); (2) reversing the order of the samples changes no table bit (sums over samples are double-double,
    rounded once) and permutes lambda; (3) evaluation counts are a per-sample property."""
    D = 1_000_000
    counts = mmsig.synth.generate(D, K, V, key=20261018)
    g0 = mmsig.synth.init_gamma(K, V)
    full = mmsig.MMCTM(K, ALPHA, counts, V=V, gamma0=g0)
    ll = full.iterate()
    sf = full.state(props=False)
    nn, nl = full.evals()
    full.close()
    assert np.isfinite(ll).all() and np.isfinite(sf["lam"]).all()
    # (1) two halves
    h = D // 2
    stats = []
    for lo, hi in ((0, h), (h, D)):
        part = [(r[lo:hi + 1] - r[lo], t[r[lo]:r[hi]], c[r[lo]:r[hi]]) for r, t, c in counts]
        m = mmsig.MMCTM(K, ALPHA, part, V=V, gamma0=g0)
        m.iterate()
        sp = m.state(props=False)
        a, b = m.evals()
        assert np.array_equal(sp["lam"], sf["lam"][lo:hi]) and np.array_equal(sp["nu"], sf["nu"][lo:hi])
        assert np.array_equal(a, nn[lo:hi]) and np.array_equal(b, nl[lo:hi])
        stats.append(sp["gamma"] - 0.1)
        m.close()
    assert rel_err(stats[0] + stats[1], sf["gamma"] - 0.1) <= 1e-13
    # (2) reversed sample order
    rev = []
    for r, t, c in counts:
        n = np.diff(r)[::-1]
        rr = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
        idx = np.concatenate([np.arange(r[d], r[d + 1]) for d in range(D - 1, D - 1001, -1)])    # spot-check construction
        tt = np.empty_like(t); cc = np.empty_like(c)
        # vectorised reversal of the rows
        starts = r[:-1][::-1]
        pos = np.repeat(starts - rr[:-1], n) + np.arange(rr[-1])
        tt[:] = t[pos]; cc[:] = c[pos]
        assert np.array_equal(tt[:len(idx)], t[idx])
        rev.append((rr, tt, cc))
    m = mmsig.MMCTM(K, ALPHA, rev, V=V, gamma0=g0)
    ll_r = m.iterate()
    sr = m.state(props=False)
    m.close()
    assert np.array_equal(sr["lam"], sf["lam"][::-1])
    for k in ("gamma", "mu", "Sigma", "invSigma", "phi"):
        assert np.array_equal(sr[k], sf[k]), k
    assert np.array_equal(ll_r, ll)

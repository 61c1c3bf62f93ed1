"""Several GPUs from ONE process (mmsig_group_*, include/mmsig.h): the samples of one fit sharded over the
members, partial sums exchanged through peer memory, every member reducing in rank order.  The result has to be
bit-identical to the one-GPU fit and to the oracle.  With one GPU the same device is listed twice (two shards,
two streams, the same exchange code); with two or more, distinct devices (NVLink peer stores)."""
import numpy as np
import pytest

import mmsig
from util import oracle_mmctm, small_synth, rel_err

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _device_sets():
    sets = [[0, 0], [0, 0, 0]]
    if _ndev() >= 2:
        sets.append([0, 1])
    if _ndev() >= 4:
        sets.append([0, 1, 2, 3])
    return sets


K, V, ALPHA = [10, 8, 6], [96, 32, 83], [0.1, 0.1, 0.1]


@pytest.mark.parametrize("devices", _device_sets() if True else [])
def test_group_fit_is_bit_identical_to_one_gpu_and_oracle(devices):
    D = 1500
    counts = small_synth(D, K, V, empty_frac=0.05)
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, ALPHA, V, counts, g0)
    one = mmsig.MMCTM(K, ALPHA, counts, V=V, gamma0=g0)
    grp = mmsig.MMCTMGroup(K, ALPHA, counts, devices, V=V, gamma0=g0)
    for _ in range(3):
        ll_o, ll_1, ll_g = o.iterate(), one.iterate(), grp.iterate()
        assert np.array_equal(ll_g, ll_1) and np.array_equal(ll_g, ll_o)
    s1, sg = one.state(), grp.state()
    for k in s1:
        assert np.array_equal(s1[k], sg[k]), k
    assert np.array_equal(sg["lam"], o.lam) and np.array_equal(sg["Sigma"], o.Sigma) and np.array_equal(sg["phi"], o.phi)
    a, b = grp.evals()
    assert np.array_equal(a, o.nev_nu) and np.array_equal(b, o.nev_lambda)
    eo, eg = o.elbo()[0], grp.calculate_elbo()[0]
    assert abs(eg - eo) <= 1e-12 * abs(eo)
    assert grp.grp.launch_count() > 0
    one.close(); grp.close()


@pytest.mark.parametrize("devices", _device_sets())
def test_group_fit_host_and_fit(devices):
    D = 2500
    counts = small_synth(D, K, V, seed=11, empty_frac=0.03)
    g0 = mmsig.synth.init_gamma(K, V)
    one = mmsig.MMCTM(K, ALPHA, counts, V=V, gamma0=g0)
    h1 = one.fit(maxiter=13, tol=1e-3, verbose=False)
    s1 = one.state()
    grp = mmsig.MMCTMGroup(K, ALPHA, small_synth(200, K, V, seed=2), devices, V=V, gamma0=g0)    # planned for another corpus
    hg, sg = grp.fit_host(counts, g0, maxiter=13, tol=1e-3)
    assert np.array_equal(h1, hg) and grp.converged == one.converged
    for k in s1:
        assert np.array_equal(s1[k], sg[k]), k
    assert grp.calculate_elbo()[0] == one.elbo
    # resident fit through the group, continuing from that state
    grp.set_state(sg["gamma"], lam=sg["lam"], nu=sg["nu"], mu=sg["mu"], Sigma=sg["Sigma"], invSigma=sg["invSigma"])
    one.set_state(s1["gamma"], lam=s1["lam"], nu=s1["nu"], mu=s1["mu"], Sigma=s1["Sigma"], invSigma=s1["invSigma"])
    assert np.array_equal(grp.fit(maxiter=3, tol=1e-9), one.fit(maxiter=3, tol=1e-9, verbose=False))
    one.close(); grp.close()


@pytest.mark.parametrize("devices", _device_sets())
def test_group_restarts_dealt_over_devices(devices):
    """config 5 (shape): restarts dealt over the members, arg-max ELBO on the host; equal to the one-handle call."""
    Kr, Vr, D, R = [7, 7], [96, 32], 400, 5
    counts = small_synth(D, Kr, Vr)
    g0s = np.random.default_rng(5).integers(1, 101, size=(R, sum(k * v for k, v in zip(Kr, Vr)))).astype(float)
    one = mmsig.MMCTM(Kr, [0.1, 0.1], counts, V=Vr, gamma0=g0s[0])
    e1, l1, n1, b1 = one.fit_restarts(g0s, maxiter=12, tol=1e-4)
    grp = mmsig.MMCTMGroup(Kr, [0.1, 0.1], counts, devices, V=Vr, gamma0=g0s[0])
    eg, lg, ng, bg = grp.fit_restarts(g0s, maxiter=12, tol=1e-4)
    assert np.array_equal(e1, eg) and np.array_equal(l1, lg) and np.array_equal(n1, ng) and b1 == bg
    s1, sg = one.state(), grp.state()
    for k in s1:
        assert np.array_equal(s1[k], sg[k]), k
    assert grp.calculate_elbo()[0] == e1[b1]
    one.close(); grp.close()


def test_group_reports_a_member_failure_without_hanging():
    counts = small_synth(300, K, V)
    g0 = mmsig.synth.init_gamma(K, V)
    grp = mmsig.MMCTMGroup(K, ALPHA, counts, [0, 0], V=V, gamma0=g0)
    bad = [(r.copy(), t.copy(), c.copy()) for r, t, c in counts]
    bad[0][1][int(bad[0][0][250])] = 1000                      # a term out of range in the second shard only
    with pytest.raises(mmsig.capi.MmsigError):
        grp.fit_host(bad, g0, maxiter=2)
    hist, _ = grp.fit_host(counts, g0, maxiter=2)              # the group recovers
    assert np.isfinite(hist).all()
    with pytest.raises(mmsig.capi.MmsigError):                 # fewer samples than members
        mmsig.MMCTMGroup(K, ALPHA, small_synth(2, K, V), [0, 0, 0], V=V, gamma0=g0)
    grp.close()

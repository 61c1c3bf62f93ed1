"""Edge cases of the MMCTM / LDA paths against the oracle (bit-exact for MMCTM)."""
import numpy as np
import pytest

import orc
import mmsig
from mmsig.counts import make_count_csr
from util import oracle_mmctm, small_synth, rel_err

pytestmark = pytest.mark.gpu


def _same(o, g, ll_o, ll_g):
    s = g.state()
    for k in ("lam", "nu", "zeta", "gamma", "mu", "Sigma", "invSigma", "Elnphi", "phi", "props"):
        assert np.array_equal(s[k], getattr(o, k), equal_nan=True), k
    assert np.array_equal(ll_g, ll_o, equal_nan=True)


def _run(K, V, counts, iters=2, seed=0, state=None):
    alpha = [0.1] * len(K)
    G = sum(k * v for k, v in zip(K, V))
    g0 = np.random.default_rng(seed).integers(1, 101, G).astype(float)
    o = oracle_mmctm(K, alpha, V, counts, g0)
    g = mmsig.MMCTM(K, alpha, counts, V=V, gamma0=g0)
    if state is not None:
        lam, nu, mu, Sig = state
        iS = np.linalg.inv(Sig)
        g.set_state(g0, lam=lam, nu=nu, mu=mu, Sigma=Sig, invSigma=iS)
        o.lam[:] = lam; o.nu[:] = nu; o.mu[:] = mu; o.Sigma[:] = Sig; o.invSigma[:] = iS
        for d in range(o.D):
            o.L.orc_mmctm_update_zeta(o.p, d)
    for _ in range(iters):
        _same(o, g, o.iterate(), g.iterate())
    eo, eg = o.elbo()[0], g.calculate_elbo()[0]
    assert (np.isnan(eo) and np.isnan(eg)) or abs(eg - eo) <= 1e-12 * abs(eo)
    g.close()


def test_single_sample():
    _run([3, 2], [9, 5], small_synth(1, [3, 2], [9, 5]))


def test_sample_with_every_modality_empty_and_one_term_vocab():
    K, V, D = [2, 1, 3], [6, 1, 4], 40
    rng = np.random.default_rng(4)
    dense = [rng.poisson(3.0, size=(v, D)) for v in V]
    for x in dense:
        x[:, :5] = 0                       # samples 0..4: nothing observed at all
    dense[1][:, 7] = 0
    _run(K, V, [make_count_csr(x) for x in dense], iters=3)


def test_modality_that_is_empty_for_all_samples():
    """N_m = 0 for every sample: the reference's ll[m] is 0/0 = NaN (src/MMCTM.jl:417); so is ours."""
    K, V, D = [2, 2], [5, 4], 30
    rng = np.random.default_rng(5)
    a = rng.poisson(4.0, size=(5, D))
    b = np.zeros((4, D), dtype=np.int64)
    counts = [make_count_csr(a), make_count_csr(b)]
    alpha = [0.1, 0.1]
    g0 = rng.integers(1, 101, 18).astype(float)
    o = oracle_mmctm(K, alpha, V, counts, g0)
    g = mmsig.MMCTM(K, alpha, counts, V=V, gamma0=g0)
    ll_o, ll_g = o.iterate(), g.iterate()
    assert np.isnan(ll_o[1]) and np.isnan(ll_g[1]) and ll_o[0] == ll_g[0]
    s = g.state()
    assert np.array_equal(s["lam"], o.lam) and np.array_equal(s["phi"], o.phi)
    g.close()


def test_huge_counts_and_nondefault_state():
    K, V, D = [4, 3], [20, 10], 120
    rng = np.random.default_rng(6)
    dense = [rng.poisson(2000.0, size=(20, D)) * rng.integers(0, 2, size=(20, D)), rng.poisson(1.5, size=(10, D))]
    dense[0][3, 10] = 2_000_000
    MK = 7
    A = rng.standard_normal((MK, MK))
    state = (rng.standard_normal((D, MK)) * 0.5, rng.uniform(0.05, 2.0, (D, MK)), rng.standard_normal(MK) * 0.3,
             A @ A.T / MK + 0.5 * np.eye(MK))
    _run(K, V, [make_count_csr(x) for x in dense], iters=3, state=state)


@pytest.mark.parametrize("seed", range(6))
def test_random_small_models(seed):
    rng = np.random.default_rng(100 + seed)
    M = int(rng.integers(1, 5))
    K = [int(rng.integers(1, 9)) for _ in range(M)]
    while sum(K) > 32:
        K[int(np.argmax(K))] -= 1
    V = [int(rng.integers(1, 130)) for _ in range(M)]
    D = int(rng.integers(2, 400))
    rates = [float(rng.choice([0.3, 5.0, 80.0, 3000.0])) for _ in range(M)]
    counts = mmsig.synth.generate(D, K, V, rates=rates, key=1000 + seed)
    _run(K, V, counts, iters=2, seed=seed)


def test_long_fit_stays_bit_exact():
    K, V, D = [5, 4], [30, 12], 250
    counts = small_synth(D, K, V, seed=21)
    g0 = mmsig.synth.init_gamma(K, V)
    o = oracle_mmctm(K, [0.1, 0.1], V, counts, g0)
    g = mmsig.MMCTM(K, [0.1, 0.1], counts, V=V, gamma0=g0)
    ho = o.fit(maxiter=60, tol=1e-9)
    hg = g.fit(maxiter=60, tol=1e-9, verbose=False)
    assert np.array_equal(hg, ho)
    eo = o.elbo()[0]
    assert abs(g.elbo - eo) <= 1e-12 * abs(eo)
    g.close()


def test_lda_edge_cases():
    rng = np.random.default_rng(8)
    dense = rng.poisson(2.0, size=(7, 50))
    dense[:, :3] = 0                      # empty documents
    csr = make_count_csr(dense)
    lam0 = rng.integers(1, 101, 3 * 7).astype(float)
    o = orc.OracleLDA(3, 0.1, 0.1, 7, csr, lam0)
    g = mmsig.LDA(3, 0.1, 0.1, csr, V=7, lambda0=lam0)
    for _ in range(3):
        a, b = o.iterate(), g.iterate()
        assert abs(a - b) <= 1e-12 * abs(a)
    s = g.state()
    assert rel_err(s["gamma"], o.gamma) <= 1e-12 and rel_err(s["lam"], o.lam) <= 1e-12
    g.close()


@pytest.mark.parametrize("K,V,D", [
    ([1], [1], 5),                       # one topic, one term
    ([32], [33], 70),                    # K at its per-modality limit, V one past a warp
    ([3, 17], [64, 65], 97),             # V on and one past a 32-term block boundary; D not a multiple of the tile
    ([2, 3, 1, 2, 3, 1, 2, 2], [7, 33, 2, 12, 40, 3, 31, 32], 45),   # 8 modalities (the limit)
    ([5, 4], [300, 17], 31),             # several warps of terms, fewer samples than one tile
    ([13, 12, 11], [40, 100, 9], 65),    # sum(K) = 36: the two-coordinates-per-lane path with the tile kernels
])
def test_awkward_shapes_of_the_tile_kernels(K, V, D):
    """Term counts around the 32-term blocks, sample counts around the 32-sample tiles, the K and M
    limits, both solve layouts: θ pass, solve, log-likelihood and ELBO against the oracle."""
    rng = np.random.default_rng(sum(V) + D)
    dense = [rng.poisson(1.5, size=(v, D)) for v in V]
    dense[0][:, D // 2] = 0                                   # an empty row in the first modality
    _run(K, V, [make_count_csr(x) for x in dense], iters=2, seed=D)

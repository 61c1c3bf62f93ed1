"""Synthetic Poisson counts from a generative MMCTM (BASELINE.json north_star).

Truth:  phi*_{mk} ~ Dirichlet(0.1 1_V);  A ~ N(0,1)^{MKxMK}, Sigma* = A A^T / MK + 0.1 I, mu* = 0;
        eta_d ~ N(mu*, Sigma*);  p_dm = softmax(eta_d[block m]);
        count_dmv ~ Poisson(R_m sum_k p_dmk phi*_mkv),  R = 3500 (SNV96) / 85 (SV32, SV48) / 300 (ID83).
RNG:    numpy Generator(Philox).  The truth uses key DATA_KEY; samples are drawn in chunks of CHUNK
        samples, chunk i from Philox(key=DATA_KEY + 1 + i), so any contiguous range of samples can be
        generated on its own (each rank generates only its shard) and the union does not depend on
        the number of ranks.  Init state: Philox(key=INIT_KEY), gamma0 = integers U{1..100}
        (src/MMCTM.jl:61), lambda0 likewise (src/LDA.jl:36).
"""
import numpy as np

from .counts import make_count_csr

DATA_KEY = 20261018
INIT_KEY = 42
CHUNK = 50_000
RATES = {96: 3500.0, 32: 85.0, 83: 300.0, 48: 85.0}


def truth(K, V, key=DATA_KEY):
    rng = np.random.Generator(np.random.Philox(key=key))
    MK = int(sum(K))
    phis = [rng.dirichlet(np.full(V[m], 0.1), size=K[m]) for m in range(len(K))]
    A = rng.standard_normal((MK, MK))
    Sig = A @ A.T / MK + 0.1 * np.eye(MK)
    return phis, Sig


def generate(D, K, V, rates=None, key=DATA_KEY, lo=0, hi=None):
    """CSR counts of samples lo:hi (default all D) of the D-sample synthetic corpus."""
    hi = D if hi is None else hi
    M = len(K)
    rates = [RATES.get(v, 300.0) for v in V] if rates is None else rates
    phis, Sig = truth(K, V, key)
    Lc = np.linalg.cholesky(Sig)
    MK = int(sum(K))
    parts = [[] for _ in range(M)]
    for ci in range(lo // CHUNK, (max(hi, 1) - 1) // CHUNK + 1):
        c0, c1 = ci * CHUNK, min((ci + 1) * CHUNK, D)
        rng = np.random.Generator(np.random.Philox(key=key + 1 + ci))
        eta = rng.standard_normal((c1 - c0, MK)) @ Lc.T
        a, b = max(lo, c0) - c0, min(hi, c1) - c0
        off = 0
        for m in range(M):
            e = eta[:, off:off + K[m]]
            e = np.exp(e - e.max(axis=1, keepdims=True))
            p = e / e.sum(axis=1, keepdims=True)
            cnt = rng.poisson(rates[m] * (p @ phis[m])).astype(np.int32)
            parts[m].append(cnt[a:b])
            off += K[m]
    return [make_count_csr(np.concatenate(parts[m], axis=0).T) for m in range(M)]


def init_gamma(K, V, key=INIT_KEY):
    """gamma0 = rand(1:100, V[m]) per topic (src/MMCTM.jl:59-63), flat [m][k][v]."""
    rng = np.random.Generator(np.random.Philox(key=key))
    return np.concatenate([rng.integers(1, 101, size=K[m] * V[m]).astype(np.float64)
                           for m in range(len(K))])


def init_lda_lambda(K, V, key=INIT_KEY):
    """lambda0 = rand(1:100, V, K) (src/LDA.jl:36), flat [k][v]."""
    rng = np.random.Generator(np.random.Philox(key=key))
    return rng.integers(1, 101, size=K * V).astype(np.float64)

"""Synthetic Poisson counts from a generative MMCTM (BASELINE.json north_star).

Data: numpy Generator(Philox(key=20261018)); init: Philox(key=42).
phi*_{mk} ~ Dirichlet(0.1); A ~ N(0,1)^{MKxMK}, Sigma* = A A^T / MK + 0.1 I, mu* = 0;
eta_d ~ N(mu*, Sigma*); p_dm = softmax(eta_d[block m]); count_dmv ~ Poisson(R_m sum_k p_dmk phi*_mkv).
"""
import numpy as np

from .counts import make_count_csr

DATA_KEY = 20261018
INIT_KEY = 42
RATES = {96: 3500.0, 32: 85.0, 83: 300.0, 48: 85.0}


def generate(D, K, V, rates=None, key=DATA_KEY, chunk=200_000):
    rng = np.random.Generator(np.random.Philox(key=key))
    M = len(K)
    MK = int(sum(K))
    rates = [RATES.get(v, 300.0) for v in V] if rates is None else rates
    phis = [rng.dirichlet(np.full(V[m], 0.1), size=K[m]) for m in range(M)]
    A = rng.standard_normal((MK, MK))
    Sig = A @ A.T / MK + 0.1 * np.eye(MK)
    Lc = np.linalg.cholesky(Sig)
    parts = [[] for _ in range(M)]
    for lo in range(0, D, chunk):
        n = min(chunk, D - lo)
        eta = rng.standard_normal((n, MK)) @ Lc.T
        off = 0
        for m in range(M):
            e = eta[:, off:off + K[m]]
            e = np.exp(e - e.max(axis=1, keepdims=True))
            p = e / e.sum(axis=1, keepdims=True)
            mean = rates[m] * (p @ phis[m])
            parts[m].append(rng.poisson(mean).astype(np.int32))
            off += K[m]
    counts = []
    for m in range(M):
        dense = np.concatenate(parts[m], axis=0)       # (D, V)
        counts.append(make_count_csr(dense.T))
    return counts


def init_gamma(K, V, key=INIT_KEY):
    """gamma0 = rand(1:100, V[m]) per topic (src/MMCTM.jl:59-63), flat [m][k][v]."""
    rng = np.random.Generator(np.random.Philox(key=key))
    return np.concatenate([rng.integers(1, 101, size=K[m] * V[m]).astype(np.float64)
                           for m in range(len(K))])


def init_lda_lambda(K, V, key=INIT_KEY):
    """lambda0 = rand(1:100, V, K) (src/LDA.jl:36), flat [k][v]."""
    rng = np.random.Generator(np.random.Philox(key=key))
    return rng.integers(1, 101, size=K * V).astype(np.float64)

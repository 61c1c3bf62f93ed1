"""Shared helpers for the parity tests."""
import numpy as np

import orc
from mmsig import synth
from mmsig.counts import make_count_csr


def oracle_mmctm(K, alpha, V, counts, gamma0, arith=orc.ARITH_DET, stop_rule=0, nthreads=8):
    return orc.OracleMMCTM(K, alpha, V, counts, gamma0, arith=arith, stop_rule=stop_rule, nthreads=nthreads)


def small_synth(D, K, V, seed=7, empty_frac=0.0):
    counts = synth.generate(D, K, V, key=seed)
    if empty_frac > 0:                         # knock out some rows entirely (empty modality rows)
        rng = np.random.default_rng(seed)
        out = []
        for (rp, t, c), v in zip(counts, V):
            dense = np.zeros((D, v), dtype=np.int64)
            for d in range(D):
                dense[d, t[rp[d]:rp[d + 1]]] = c[rp[d]:rp[d + 1]]
            dense[rng.random(D) < empty_frac] = 0
            out.append(make_count_csr(dense.T))
        counts = out
    return counts


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def norm_err(a, b):
    """max |a-b| / max |b| : the right yardstick for vectors with entries near 0 (μ, off-diagonal Σ)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if a.size else 0.0

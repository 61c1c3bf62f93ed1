#!/usr/bin/env python
"""Timings of the other BASELINE.json configs (parity-test cases, not bench lines): config 2
(LDA K=20), config 3 (CTM K=10), config 5 (restarts of MMCTM([7,7]) on 100k samples), config 1
(brca-eu fit).  Usage (under gpurun): python profiles/bench_configs.py [D]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import mmsig  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
out = {}


def timed(fn, n):
    fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t) / n * 1e3


# config 2: LDA(20, 0.1, 0.1), D x 96
csr = mmsig.synth.generate(D, [20], [96])[0]
nnz = int(csr[0][-1])
m = mmsig.LDA(20, 0.1, 0.1, csr, V=96, lambda0=mmsig.synth.init_lda_lambda(20, 96), profile=True)
for _ in range(3):
    m.iterate()
m.h.kernel_times(reset=True)
ms = timed(m.iterate, 10)
kt = m.h.kernel_times(reset=True)
alg = 16.0 * nnz + 24.0 * 20 * D
out["config2_lda_k20"] = {"D": D, "nnz_per_sample": nnz / D, "ms_per_iteration": ms, "iterations_per_s": 1e3 / ms,
                          "algorithmic_GBs": alg / (ms * 1e-3) / 1e9,
                          "kernels_ms": {k: v[0] / max(v[1], 1) * (v[1] / 11.0) for k, v in kt.items()}}
m.close()

# config 3: CTM = MMCTM([10]), D x 96
m = mmsig.MMCTM([10], [0.1], [csr], V=[96], gamma0=mmsig.synth.init_gamma([10], [96]), profile=True)
for _ in range(3):
    m.iterate()
m.h.kernel_times(reset=True)
ms = timed(m.iterate, 5)
kt = m.h.kernel_times(reset=True)
alg = 16.0 * nnz + D * (40.0 * 10 + 24.0)
out["config3_ctm_k10"] = {"D": D, "ms_per_iteration": ms, "iterations_per_s": 1e3 / ms,
                          "algorithmic_GBs": alg / (ms * 1e-3) / 1e9,
                          "kernels_ms": {k: v[0] / max(v[1], 1) * (v[1] / 6.0) for k, v in kt.items()}}
m.close()
del csr

# config 5 (one GPU's share): 8 restarts of MMCTM([7,7]) on 100k samples, V = [96, 32]
D5 = min(D, 100_000)
c5 = mmsig.synth.generate(D5, [7, 7], [96, 32])
rng = np.random.Generator(np.random.Philox(key=7))
g0s = rng.integers(1, 101, size=(8, 7 * 96 + 7 * 32)).astype(float)
m = mmsig.MMCTM([7, 7], [0.1, 0.1], c5, V=[96, 32], gamma0=g0s[0])
t = time.perf_counter()
elbo, ll, nit, best = m.fit_restarts(g0s, maxiter=30, tol=1e-4)
dt = time.perf_counter() - t
out["config5_restarts_per_gpu"] = {"D": D5, "restarts": 8, "iterations": nit.tolist(), "seconds": dt,
                                   "ms_per_iteration": dt * 1e3 / float(nit.sum()), "best": int(best),
                                   "elbo": elbo.tolist()}
m.close()

# config 1: brca-eu, MMCTM([7,7]), fit!(tol=1e-5)
z = np.load(os.path.join(ROOT, "tests", "golden", "brca_eu_counts.npz"))
brca = [(z["rowptr0"], z["term0"], z["count0"]), (z["rowptr1"], z["term1"], z["count1"])]
m = mmsig.MMCTM([7, 7], [0.1, 0.1], brca, V=[96, 48], gamma0=mmsig.synth.init_gamma([7, 7], [96, 48]))
t = time.perf_counter()
hist = m.fit(maxiter=100, tol=1e-5, verbose=False)
dt = time.perf_counter() - t
out["config1_brca_fit"] = {"D": 560, "iterations": len(hist), "seconds": dt, "ms_per_iteration": dt * 1e3 / len(hist),
                           "converged": m.converged, "elbo": m.elbo, "ll": hist[-1].tolist()}
m.close()
print(json.dumps(out, indent=1))

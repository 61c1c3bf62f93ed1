"""Generates tests/golden/config1_det.json: the pinned-arithmetic oracle's results on BASELINE
config 1 (MMCTM([7,7],[0.1,0.1]) on the bundled brca-eu counts, gamma0 from Philox(key=42)) and
on a small config-4-shaped synthetic corpus.  The reference's own tests do not pin these numbers
(SURVEY 4); the fixture pins OUR specification, so that neither the oracle nor the CUDA path can
drift silently.  Re-run after any deliberate change of the DET specification:
    python tests/golden/make_config1_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmsig  # noqa: E402
import orc  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def run(name, K, V, counts, iters, out):
    g0 = mmsig.synth.init_gamma(K, V)
    o = orc.OracleMMCTM(K, [0.1] * len(K), V, counts, g0, arith=orc.ARITH_DET, nthreads=os.cpu_count() or 1)
    hist = o.fit(maxiter=iters, tol=1e-5)
    elbo, terms = o.elbo()
    out[name] = {"K": K, "V": V, "iterations": int(len(hist)), "converged": bool(o.converged),
                 "ll_history_hex": [[float(x).hex() for x in r] for r in hist],
                 "elbo": elbo, "elbo_terms": terms.tolist(),
                 "sha256": {k: digest(getattr(o, k)) for k in ("lam", "nu", "zeta", "gamma", "phi", "mu", "Sigma", "invSigma", "props")},
                 "evals_last_iteration": {"nu_sum": int(o.nev_nu.sum()), "lambda_sum": int(o.nev_lambda.sum())}}


out = {}
z = np.load(os.path.join(ROOT, "tests", "golden", "brca_eu_counts.npz"))
brca = [(z["rowptr0"], z["term0"], z["count0"]), (z["rowptr1"], z["term1"], z["count1"])]
run("config1_brca_eu", [7, 7], [96, 48], brca, 25, out)
K, V = [10, 8, 6], [96, 32, 83]
run("config4_shape_D2000", K, V, mmsig.synth.generate(2000, K, V, key=mmsig.synth.DATA_KEY), 6, out)
path = os.path.join(ROOT, "tests", "golden", "config1_det.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path, {k: (v["iterations"], v["elbo"]) for k, v in out.items()})

"""tests/golden/immctm_known_answers.json: the known answers of the reference's own IMMCTM tests
(/root/reference/test/immctm.jl), transcribed as closed-form expressions and evaluated with
scipy.special.digamma (the reference evaluates the same expressions with SpecialFunctions.digamma).
Run here (the reference tree is not needed: the expressions are restated below with their lines)."""
import json
import os

import numpy as np
from scipy.special import digamma as psi

out = {
    # test/immctm.jl:6-50 -- the toy corpus (1-based terms and feature values as in Julia)
    "K": [2, 3], "alpha": [0.1, 0.1],
    "features": [[[1, 1], [1, 2], [2, 1], [2, 2]], [[1, 1], [1, 2], [2, 1], [2, 2]]],
    "X": [[[[1, 5], [2, 8]], [[1, 2], [2, 5]]], [[[3, 4], [4, 9]], [[3, 4], [4, 6]]]],
    # :52-62 constructor
    "ctor": {"N": [[13, 7], [13, 10]], "I": [2, 2], "J": [[2, 2], [2, 2]], "V": [4, 4]},
}
# :181-224 update_θ!
lam = [[1, 2, 3, 4, 1], [2, 3, 1, 4, 2]]
gam = [[[[0.1, 0.2], [0.1, 1.0]], [[0.1, 0.1], [1.0, 1.0]]],
       [[[0.5, 0.5], [1.0, 1.5]], [[1.0, 2.0], [2.0, 3.0]], [[1.0, 5.0], [5.0, 2.0]]]]
t1 = np.empty((2, 2))
t1[0, 0] = np.exp(1 + psi(0.1) - psi(0.3) + psi(0.1) - psi(1.1))
t1[1, 0] = np.exp(2 + psi(0.1) - psi(0.2) + psi(1.0) - psi(2.0))
t1[0, 1] = np.exp(1 + psi(0.1) - psi(0.3) + psi(1.0) - psi(1.1))
t1[1, 1] = np.exp(2 + psi(0.1) - psi(0.2) + psi(1.0) - psi(2.0))
t1 /= t1.sum(axis=0)
t2 = np.empty((3, 2))
t2[0, 0] = np.exp(1 + psi(0.5) - psi(1.0) + psi(1.0) - psi(2.5))
t2[1, 0] = np.exp(4 + psi(2.0) - psi(3.0) + psi(2.0) - psi(5.0))
t2[2, 0] = np.exp(2 + psi(5.0) - psi(6.0) + psi(5.0) - psi(7.0))
t2[0, 1] = np.exp(1 + psi(0.5) - psi(1.0) + psi(1.5) - psi(2.5))
t2[1, 1] = np.exp(4 + psi(2.0) - psi(3.0) + psi(3.0) - psi(5.0))
t2[2, 1] = np.exp(2 + psi(5.0) - psi(6.0) + psi(2.0) - psi(7.0))
t2 /= t2.sum(axis=0)
out["update_theta"] = {"lambda": lam, "gamma": gam, "expected_theta_d1_m1": t1.tolist(), "expected_theta_d2_m2": t2.tolist()}
# :251-262 update_γ!
out["update_gamma"] = {"theta_d1_m1": [[0.4, 0.1], [0.6, 0.9]], "theta_d2_m1": [[0.3, 0.5], [0.7, 0.5]],
                       "expected_m1_k1_i1": [0.1 + 5 * 0.4 + 8 * 0.1, 0.1 + 4 * 0.3 + 9 * 0.5],
                       "expected_m1_k1_i2": [0.1 + 5 * 0.4 + 4 * 0.3, 0.1 + 8 * 0.1 + 9 * 0.5]}
# :264-271 update_Elnϕ!
out["update_Elnphi"] = {"gamma_m1_k1_i1": [1, 2], "expected_first": float(psi(1) - psi(3))}
# :350-386 calculate_modality_loglikelihood
eta = [[1.0, 2.0], [2.0, 3.0]]
th = [np.exp(e) / np.exp(e).sum() for e in eta]
g = [[[0.1, 0.2], [0.1, 1.0]], [[0.1, 0.1], [1.0, 1.0]]]
ph = [[np.asarray(g[k][i]) / sum(g[k][i]) for i in range(2)] for k in range(2)]
s = (5 * np.log(th[0][0] * ph[0][0][0] * ph[0][1][0] + th[0][1] * ph[1][0][0] * ph[1][1][0]) +
     8 * np.log(th[0][0] * ph[0][0][0] * ph[0][1][1] + th[0][1] * ph[1][0][0] * ph[1][1][1]) +
     4 * np.log(th[1][0] * ph[0][0][1] * ph[0][1][0] + th[1][1] * ph[1][0][1] * ph[1][1][0]) +
     9 * np.log(th[1][0] * ph[0][0][1] * ph[0][1][1] + th[1][1] * ph[1][0][1] * ph[1][1][1]))
out["loglikelihood"] = {"eta": eta, "gamma_m1": g, "expected_m1": float(s / 26.0)}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "immctm_known_answers.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)

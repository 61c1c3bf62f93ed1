"""tests/golden/ilda_known_answers.json: the known answers of the reference's own ILDA tests
(/root/reference/test/ilda.jl), transcribed as closed-form expressions (scipy digamma)."""
import json
import os

import numpy as np
from scipy.special import digamma as psi

out = {"K": 2, "alpha": 0.1, "eta": 0.1,                                   # test/ilda.jl:4-22
       "features": [[1, 1], [1, 2], [2, 1], [2, 2]], "X": [[[1, 5], [2, 8]], [[3, 2], [4, 5]]],
       "ctor": {"I": 2, "J": [2, 2]}}                                        # :24-34
# :52-92 update_ϕ!   (Elnβ[i] is J_i x K: Elnβ[i][j, k])
Et = np.array([[0.5, -1.1], [2.3, -0.7]])
Eb = [np.array([[-0.2, -0.9], [-1.1, 0.3]]), np.array([[0.5, 0.1], [-0.1, -0.4]])]
p1 = np.empty((2, 2))
p1[0, 0] = np.exp(Et[0, 0] + Eb[0][0, 0] + Eb[1][0, 0]); p1[0, 1] = np.exp(Et[0, 0] + Eb[0][0, 0] + Eb[1][1, 0])
p1[1, 0] = np.exp(Et[1, 0] + Eb[0][0, 1] + Eb[1][0, 1]); p1[1, 1] = np.exp(Et[1, 0] + Eb[0][0, 1] + Eb[1][1, 1])
p1 /= p1.sum(axis=0)
p2 = np.empty((2, 2))
p2[0, 0] = np.exp(Et[0, 1] + Eb[0][1, 0] + Eb[1][0, 0]); p2[0, 1] = np.exp(Et[0, 1] + Eb[0][1, 0] + Eb[1][1, 0])
p2[1, 0] = np.exp(Et[1, 1] + Eb[0][1, 1] + Eb[1][0, 1]); p2[1, 1] = np.exp(Et[1, 1] + Eb[0][1, 1] + Eb[1][1, 1])
p2 /= p2.sum(axis=0)
out["update_phi"] = {"Elntheta": Et.tolist(), "Elnbeta": [e.tolist() for e in Eb], "expected_d1": p1.tolist(), "expected_d2": p2.tolist()}
# :94-111 update_γ!
ph = np.array([[0.4, 0.2], [0.6, 0.8]])
g = np.array([0.1 + ph[0, 0] * 5 + ph[0, 1] * 8, 0.1 + ph[1, 0] * 5 + ph[1, 1] * 8])
out["update_gamma"] = {"phi_d1": ph.tolist(), "expected_gamma_d1": g.tolist(), "expected_Elntheta_d1": (psi(g) - psi(g.sum())).tolist()}
# :113-158 update_λ!   (λ[i] is J_i x K)
eta = [0.1, 0.2]
P = [np.array([[0.4, 0.2], [0.6, 0.8]]), np.array([[0.1, 0.6], [0.9, 0.4]])]
X = [[5, 8], [2, 5]]
l1 = np.array([[eta[0] + P[0][0, 0] * X[0][0] + P[0][0, 1] * X[0][1], eta[0] + P[0][1, 0] * X[0][0] + P[0][1, 1] * X[0][1]],
               [eta[0] + P[1][0, 0] * X[1][0] + P[1][0, 1] * X[1][1], eta[0] + P[1][1, 0] * X[1][0] + P[1][1, 1] * X[1][1]]])
l2 = np.array([[eta[1] + P[0][0, 0] * X[0][0] + P[1][0, 0] * X[1][0], eta[1] + P[0][1, 0] * X[0][0] + P[1][1, 0] * X[1][0]],
               [eta[1] + P[0][0, 1] * X[0][1] + P[1][0, 1] * X[1][1], eta[1] + P[0][1, 1] * X[0][1] + P[1][1, 1] * X[1][1]]])
out["update_lambda"] = {"eta": eta, "phi": [p.tolist() for p in P],
                        "expected_lambda": [l1.tolist(), l2.tolist()],
                        "expected_Elnbeta": [(psi(l) - psi(l.sum(axis=0))).tolist() for l in (l1, l2)]}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ilda_known_answers.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)

"""Generates tests/golden/reference_known_answers.json.

The reference (Julia + NLopt) cannot run in the build container, so the golden vectors are the
KNOWN ANSWERS OF THE REFERENCE'S OWN TESTS: each expected expression in /root/reference/test/
{mmctm,lda,common}.jl is hand-expanded there; this script evaluates those same expansions with
mpmath at 50 digits (independent of the oracle's arithmetic) and stores inputs + expected
outputs.  Re-run:  python tests/golden/make_golden.py
"""
import json
import os

import mpmath as mp

mp.mp.dps = 50
F = float
psi = mp.digamma
lg = mp.loggamma
e = mp.e

out = {}

# toy corpus, test/mmctm.jl:4-33 (1-based terms as in the reference)
K = [2, 3]
alpha = [0.1, 0.1]
X = [[[[1, 5], [2, 8]], [[1, 2], [2, 5]]],
     [[[3, 4], [4, 9]], [[3, 4], [4, 6]]]]
out["mmctm_toy"] = {"K": K, "alpha": alpha, "X": X, "src": "test/mmctm.jl:4-33"}

# constructor, test/mmctm.jl:35-57
out["ctor"] = {"N": [[13, 7], [13, 10]], "V": [4, 4], "D": 2, "M": 2, "MK": 5,
               "src": "test/mmctm.jl:35-57"}

# calculate_Ndivzeta, test/mmctm.jl:59-72
zeta = [[2, 3], [4, 5]]
out["Ndivzeta"] = {"zeta": zeta, "expected_d1": [13 / 2, 13 / 2, 7 / 3, 7 / 3, 7 / 3],
                   "src": "test/mmctm.jl:59-72"}

# calculate_sumtheta, test/mmctm.jl:74-90
th1 = [[0.4, 0.1], [0.6, 0.9]]
th2 = [[0.3, 0.4], [0.3, 0.5], [0.4, 0.1]]
out["sumtheta"] = {
    "theta_d1": [th1, th2],
    "expected_d1": [F(5 * mp.mpf(th1[0][0]) + 8 * mp.mpf(th1[0][1])),
                    F(5 * mp.mpf(th1[1][0]) + 8 * mp.mpf(th1[1][1])),
                    F(2 * mp.mpf(th2[0][0]) + 5 * mp.mpf(th2[0][1])),
                    F(2 * mp.mpf(th2[1][0]) + 5 * mp.mpf(th2[1][1])),
                    F(2 * mp.mpf(th2[2][0]) + 5 * mp.mpf(th2[2][1]))],
    "src": "test/mmctm.jl:74-90"}

# lambda_objective, test/common.jl:35-97
mu = [1, 1, 2, 2, 1]
lam = [1, 2, 3, 4, 1]
nu = [1, 1, 1, 2, 1]
zt = [2, 1]
N1 = [13, 7]


def lam_obj():
    diff = [mp.mpf(lam[j] - mu[j]) for j in range(5)]
    quad = sum(d * d for d in diff)          # invSigma = I
    L = (-quad / 2
         + 5 * (mp.mpf(th1[0][0]) * lam[0] + mp.mpf(th1[1][0]) * lam[1])
         + 8 * (mp.mpf(th1[0][1]) * lam[0] + mp.mpf(th1[1][1]) * lam[1])
         + 2 * (mp.mpf(th2[0][0]) * lam[2] + mp.mpf(th2[1][0]) * lam[3] + mp.mpf(th2[2][0]) * lam[4])
         + 5 * (mp.mpf(th2[0][1]) * lam[2] + mp.mpf(th2[1][1]) * lam[3] + mp.mpf(th2[2][1]) * lam[4])
         - mp.mpf(13) / zt[0] * (mp.exp(lam[0] + mp.mpf(nu[0]) / 2) + mp.exp(lam[1] + mp.mpf(nu[1]) / 2))
         - mp.mpf(7) / zt[1] * (mp.exp(lam[2] + mp.mpf(nu[2]) / 2) + mp.exp(lam[3] + mp.mpf(nu[3]) / 2)
                                + mp.exp(lam[4] + mp.mpf(nu[4]) / 2)))
    st = [5 * mp.mpf(th1[0][0]) + 8 * mp.mpf(th1[0][1]), 5 * mp.mpf(th1[1][0]) + 8 * mp.mpf(th1[1][1]),
          2 * mp.mpf(th2[0][0]) + 5 * mp.mpf(th2[0][1]), 2 * mp.mpf(th2[1][0]) + 5 * mp.mpf(th2[1][1]),
          2 * mp.mpf(th2[2][0]) + 5 * mp.mpf(th2[2][1])]
    c = [mp.mpf(13) / zt[0]] * 2 + [mp.mpf(7) / zt[1]] * 3
    g = [-diff[j] + st[j] - c[j] * mp.exp(lam[j] + mp.mpf(nu[j]) / 2) for j in range(5)]
    return F(L), [F(x) for x in g]


L, g = lam_obj()
out["lambda_objective"] = {"mu": mu, "lambda": lam, "nu": nu, "zeta": zt, "theta_d1": [th1, th2],
                           "expected_value": L, "expected_grad": g, "src": "test/common.jl:35-97"}


# nu_objective, test/mmctm.jl:103-148
def nu_obj():
    c = [mp.mpf(13) / zt[0]] * 2 + [mp.mpf(7) / zt[1]] * 3
    L = (-sum(mp.mpf(v) for v in nu) / 2
         - sum(c[j] * mp.exp(lam[j] + mp.mpf(nu[j]) / 2) for j in range(5))
         + sum(mp.log(v) for v in nu) / 2)
    g = [-mp.mpf(1) / 2 - c[j] / 2 * mp.exp(lam[j] + mp.mpf(nu[j]) / 2) + 1 / (2 * mp.mpf(nu[j]))
         for j in range(5)]
    return F(L), [F(x) for x in g]


L, g = nu_obj()
out["nu_objective"] = {"mu": mu, "lambda": lam, "nu": nu, "zeta": zt,
                       "expected_value": L, "expected_grad": g, "src": "test/mmctm.jl:103-148"}

# update_zeta, test/mmctm.jl:158-166
out["update_zeta"] = {"lambda": [[1, 2, 3, 4, 1], [2, 3, 1, 4, 2]], "nu": [[1, 1, 1, 2, 1], [1, 3, 1, 2, 1]],
                      "expected_d1": [F(mp.exp(1.5) + mp.exp(2.5)), F(mp.exp(3.5) + mp.exp(5) + mp.exp(1.5))],
                      "src": "test/mmctm.jl:158-166"}

# update_theta, test/mmctm.jl:168-209
gam = [[[1, 2, 2, 6], [2, 3, 1, 2]], [[1, 2, 3, 4], [2, 1, 2, 6], [1, 1, 3, 1]]]


def norm_cols(t):
    K_, W = len(t), len(t[0])
    for w in range(W):
        s = sum(t[k][w] for k in range(K_))
        for k in range(K_):
            t[k][w] = t[k][w] / s
    return [[F(x) for x in row] for row in t]


t11 = [[mp.exp(1 + psi(1) - psi(11)), mp.exp(1 + psi(2) - psi(11))],
       [mp.exp(2 + psi(2) - psi(8)), mp.exp(2 + psi(3) - psi(8))]]
t22 = [[mp.exp(1 + psi(3) - psi(10)), mp.exp(1 + psi(4) - psi(10))],
       [mp.exp(4 + psi(2) - psi(11)), mp.exp(4 + psi(6) - psi(11))],
       [mp.exp(2 + psi(3) - psi(6)), mp.exp(2 + psi(1) - psi(6))]]
out["update_theta"] = {"lambda": [[1, 2, 3, 4, 1], [2, 3, 1, 4, 2]], "gamma": gam,
                       "expected_theta_d1_m1": norm_cols(t11), "expected_theta_d2_m2": norm_cols(t22),
                       "src": "test/mmctm.jl:168-209"}

# update_mu, test/mmctm.jl:211-218
out["update_mu"] = {"lambda": [[1, 2, 3, 4, 1], [2, 3, 1, 4, 2]], "expected": [1.5, 2.5, 2.0, 4.0, 1.5],
                    "src": "test/mmctm.jl:211-218"}

# update_Sigma, test/mmctm.jl:220-236
lam2 = [[1, 2, 3, 4, 1], [2, 3, 1, 4, 2]]
nu2 = [[1, 1, 1, 2, 1], [1, 3, 1, 2, 1]]
mu2 = [1, 1, 2, 2, 1]
S = mp.zeros(5)
for d in range(2):
    diff = [lam2[d][j] - mu2[j] for j in range(5)]
    for i in range(5):
        S[i, i] += nu2[d][i]
        for j in range(5):
            S[i, j] += diff[i] * diff[j]
S = S / 2
Si = S ** -1
out["update_Sigma"] = {"lambda": lam2, "nu": nu2, "mu": mu2,
                       "expected_Sigma": [[F(S[i, j]) for j in range(5)] for i in range(5)],
                       "expected_invSigma": [[F(Si[i, j]) for j in range(5)] for i in range(5)],
                       "src": "test/mmctm.jl:220-236"}

# update_gamma, test/mmctm.jl:238-257
thg = {"d1m1": [[0.4, 0.1], [0.6, 0.9]], "d2m1": [[0.3, 0.5], [0.7, 0.5]],
       "d1m2": [[0.2, 0.6], [0.7, 0.3], [0.1, 0.1]], "d2m2": [[0.1, 0.3], [0.7, 0.5], [0.2, 0.2]]}
m = mp.mpf
out["update_gamma"] = {
    "theta": thg,
    "expected": [
        [[F(m("0.1") + 5 * m("0.4")), F(m("0.1") + 8 * m("0.1")), F(m("0.1") + 4 * m("0.3")), F(m("0.1") + 9 * m("0.5"))],
         [F(m("0.1") + 5 * m("0.6")), F(m("0.1") + 8 * m("0.9")), F(m("0.1") + 4 * m("0.7")), F(m("0.1") + 9 * m("0.5"))]],
        [[F(m("0.1") + 2 * m("0.2")), F(m("0.1") + 5 * m("0.6")), F(m("0.1") + 4 * m("0.1")), F(m("0.1") + 6 * m("0.3"))],
         [F(m("0.1") + 2 * m("0.7")), F(m("0.1") + 5 * m("0.3")), F(m("0.1") + 4 * m("0.7")), F(m("0.1") + 6 * m("0.5"))],
         [F(m("0.1") + 2 * m("0.1")), F(m("0.1") + 5 * m("0.1")), F(m("0.1") + 4 * m("0.2")), F(m("0.1") + 6 * m("0.2"))]]],
    "src": "test/mmctm.jl:238-257"}

# update_Elnphi, test/mmctm.jl:259-266
out["update_Elnphi"] = {"gamma_m1_k1": [1, 2, 1, 3], "expected_first": F(psi(1) - psi(7)),
                        "src": "test/mmctm.jl:259-266"}

# alpha_objective, test/mmctm.jl:268-277 (sum_Elnphi is an input here; value + gradient formula)
sE = mp.mpf("-12.5")
a = mp.mpf("0.1")
out["alpha_objective"] = {"alpha": 0.1, "sum_Elnphi": -12.5, "K": 2, "V": 4,
                          "expected_value": F(2 * (lg(4 * a) - 4 * lg(a)) + a * sE),
                          "expected_grad": F(4 * 2 * (psi(4 * a) - psi(a)) + sE),
                          "src": "test/mmctm.jl:268-277"}

# loglikelihoods, test/mmctm.jl:349-388
eta = [[1.0, 2.0], [2.0, 3.0]]
props = [[mp.exp(x) / sum(mp.exp(y) for y in r) for x in r] for r in eta]
gl = [[1, 2, 1, 3], [1, 1, 2, 4]]
phi = [[mp.mpf(x) / sum(r) for x in r] for r in gl]
sll = [5 * mp.log(props[0][0] * phi[0][0] + props[0][1] * phi[1][0]) + 8 * mp.log(props[0][0] * phi[0][1] + props[0][1] * phi[1][1]),
       4 * mp.log(props[1][0] * phi[0][2] + props[1][1] * phi[1][2]) + 9 * mp.log(props[1][0] * phi[0][3] + props[1][1] * phi[1][3])]
out["loglikelihoods"] = {"eta": eta, "gamma_m1": gl,
                         "expected_doc1_m1": F(sll[0] / 13), "expected_m1": F((sll[0] + sll[1]) / 26),
                         "src": "test/mmctm.jl:349-388"}

# ---- LDA, test/lda.jl
Xl = [[[1, 5], [2, 8]], [[1, 2], [2, 5]]]
out["lda_toy"] = {"K": 2, "alpha": 0.1, "eta": 0.1, "X": Xl, "N": [13, 7], "V": 2, "src": "test/lda.jl:4-36"}
Elnth = [[0.5, -1.1], [2.3, -0.7]]     # [k][d]
Elnb = [[-0.2, -0.9], [-1.1, 0.3]]     # [v][k]
ph = [[mp.exp(m(str(Elnth[0][0])) + m(str(Elnb[0][0]))), mp.exp(m(str(Elnth[0][0])) + m(str(Elnb[1][0])))],
      [mp.exp(m(str(Elnth[1][0])) + m(str(Elnb[0][1]))), mp.exp(m(str(Elnth[1][0])) + m(str(Elnb[1][1])))]]
out["lda_update_phi"] = {"Elntheta": Elnth, "Elnbeta": Elnb, "expected_phi_d1": norm_cols(ph),
                         "src": "test/lda.jl:38-61"}
phg = [[0.4, 0.2], [0.6, 0.8]]
g1 = m("0.1") + m("0.4") * 5 + m("0.2") * 8
g2 = m("0.1") + m("0.6") * 5 + m("0.8") * 8
out["lda_update_gamma"] = {"phi_d1": phg, "expected_gamma_d1": [F(g1), F(g2)],
                           "expected_Elntheta_d1": [F(psi(g1) - psi(g1 + g2)), F(psi(g2) - psi(g1 + g2))],
                           "src": "test/lda.jl:63-80"}
phl = [[[0.4, 0.2], [0.6, 0.8]], [[0.1, 0.6], [0.9, 0.4]]]    # [d][k][w]
l = [[m("0.1") + m("0.4") * 5 + m("0.1") * 2, m("0.1") + m("0.6") * 5 + m("0.9") * 2],
     [m("0.1") + m("0.2") * 8 + m("0.6") * 5, m("0.1") + m("0.8") * 8 + m("0.4") * 5]]   # [v][k]
Eb = [[psi(l[0][0]) - psi(l[0][0] + l[1][0]), psi(l[0][1]) - psi(l[0][1] + l[1][1])],
      [psi(l[1][0]) - psi(l[0][0] + l[1][0]), psi(l[1][1]) - psi(l[0][1] + l[1][1])]]
out["lda_update_lambda"] = {"phi": phl, "expected_lambda_vk": [[F(x) for x in r] for r in l],
                            "expected_Elnbeta_vk": [[F(x) for x in r] for r in Eb],
                            "src": "test/lda.jl:82-103"}

# special functions the tests lean on (digamma / lgamma known values)
out["special"] = {"digamma": {str(x): F(psi(m(str(x)))) for x in [0.1, 0.5, 1, 1.4616321449683623, 2, 3, 6.9, 7, 7.1, 11, 100.5, 1e6]},
                  "lgamma": {str(x): F(lg(m(str(x)))) for x in [0.1, 0.4, 0.5, 1, 2.5, 9.6, 100, 3500.1]}}

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_known_answers.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)

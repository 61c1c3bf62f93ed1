"""The LD_MMA restatement (oracle/mmsig_oracle.c; NLopt itself cannot run here and no
reference test pins its output: "parity unpinned") checked as an OPTIMISER: on problems with a known
or independently computed (scipy) solution it has to stop within its x-tolerance of the minimiser,
respect bounds, never increase the objective, and do so under both x-tolerance rules and both
arithmetic modes."""
import ctypes as C

import numpy as np
import pytest
from scipy import optimize

import orc

L = orc.lib()
INF = float("inf")


def mma(f, x0, lb=None, ub=None, xtol_rel=1e-4, xtol_abs=1e-4, stop_rule=orc.STOP_NLOPT27, arith=orc.ARITH_LITERAL):
    n = len(x0)
    trace = []

    def cb(nn, xp, gp, _):
        x = np.ctypeslib.as_array(xp, (nn,)).copy()
        v, g = f(x)
        if gp:
            np.ctypeslib.as_array(gp, (nn,))[:] = g
        trace.append((x, v))
        return v

    x = np.array(x0, dtype=np.float64)
    lo = np.full(n, -INF) if lb is None else np.asarray(lb, float)
    hi = np.full(n, INF) if ub is None else np.asarray(ub, float)
    minf, nouter = C.c_double(), C.c_int()
    nev = L.orc_mma_minimize(n, orc.ORC_FUNC(cb), None, orc._dp(lo), orc._dp(hi), orc._dp(x), C.byref(minf), xtol_rel, xtol_abs,
                             stop_rule, arith, C.byref(nouter))
    return x, minf.value, nev, nouter.value, trace


RULES = [orc.STOP_NLOPT27, orc.STOP_NLOPT26]
ARITHS = [orc.ARITH_LITERAL, orc.ARITH_DET]


@pytest.mark.parametrize("rule", RULES)
@pytest.mark.parametrize("arith", ARITHS)
def test_convex_quadratic(rule, arith):
    rng = np.random.default_rng(3)
    A = rng.normal(size=(6, 6))
    Q = A @ A.T + 0.5 * np.eye(6)
    b = rng.normal(size=6)
    xs = np.linalg.solve(Q, b)
    x, fmin, nev, nouter, tr = mma(lambda x: (0.5 * x @ Q @ x - b @ x, Q @ x - b), np.zeros(6), stop_rule=rule, arith=arith)
    assert nev == len(tr) and 1 <= nouter < nev
    assert np.max(np.abs(x - xs)) <= 5e-3                  # stops on the x-tolerance, not at the optimum
    assert fmin <= tr[0][1] and abs(fmin - (0.5 * x @ Q @ x - b @ x)) <= 1e-12 * max(1, abs(fmin))
    # the returned point is the best one evaluated
    assert fmin == min(v for _, v in tr)


@pytest.mark.parametrize("rule", RULES)
def test_lower_bound_is_respected_and_active(rule):
    # minimise (x0 + 1)^2 + (x1 - 2)^2 with x >= 0.5: solution (0.5, 2)
    f = lambda x: ((x[0] + 1) ** 2 + (x[1] - 2) ** 2, np.array([2 * (x[0] + 1), 2 * (x[1] - 2)]))
    x, fmin, nev, _, tr = mma(f, [3.0, 3.0], lb=[0.5, 0.5], stop_rule=rule)
    assert all(np.all(p >= 0.5) for p, _ in tr)
    assert abs(x[0] - 0.5) <= 1e-3 and abs(x[1] - 2) <= 5e-3


@pytest.mark.parametrize("arith", ARITHS)
def test_nu_and_lambda_objectives_against_scipy(arith):
    """The two per-sample problems of the E-step (src/common.jl:11-36) on a random sample: the MMA stop is within
    its tolerance of the maximiser scipy finds, and its objective is no worse than the start's."""
    rng = np.random.default_rng(11)
    MK = 8
    A = rng.normal(size=(MK, MK))
    invS = np.linalg.inv(A @ A.T / MK + 0.3 * np.eye(MK))
    mu = rng.normal(size=MK) * 0.3
    lam0 = rng.normal(size=MK) * 0.5
    nu0 = np.ones(MK)
    Ndz = np.repeat([40.0, 15.0], 4)
    sth = rng.uniform(1, 30, MK)

    def neg_lam(x, nu):
        g = np.zeros(MK)
        v = L.orc_lambda_objective(MK, orc._dp(np.ascontiguousarray(x)), orc._dp(g), orc._dp(nu), orc._dp(Ndz), orc._dp(sth),
                                   orc._dp(mu), orc._dp(np.ascontiguousarray(invS)), arith)
        return -v, -g

    def neg_nu(x, lam):
        g = np.zeros(MK)
        v = L.orc_nu_objective(MK, orc._dp(np.ascontiguousarray(x)), orc._dp(g), orc._dp(lam), orc._dp(Ndz), orc._dp(mu),
                               orc._dp(np.ascontiguousarray(invS)), arith)
        return -v, -g

    # ν: bounded below at 1e-7 (src/MMCTM.jl:160), start = current ν
    x, fmin, nev, _, tr = mma(lambda x: neg_nu(x, lam0), nu0, lb=np.full(MK, 1e-7), arith=arith)
    ref = optimize.minimize(lambda x: neg_nu(x, lam0), nu0, jac=True, method="L-BFGS-B", bounds=[(1e-7, None)] * MK,
                            options=dict(ftol=1e-15, gtol=1e-10))
    assert fmin <= tr[0][1] and np.all(x > 0)
    assert fmin - ref.fun <= 1e-3 * abs(ref.fun) + 1e-3            # MMA stops early (SURVEY finding 4), never beyond
    assert ref.fun <= fmin + 1e-9
    # λ: unbounded
    x, fmin, nev, _, tr = mma(lambda x: neg_lam(x, nu0), lam0, arith=arith)
    ref = optimize.minimize(lambda x: neg_lam(x, nu0), lam0, jac=True, method="BFGS", options=dict(gtol=1e-9))
    assert fmin <= tr[0][1]
    assert np.max(np.abs(x - ref.x)) <= 5e-3
    assert 0 <= fmin - ref.fun <= 1e-5 * abs(ref.fun) + 1e-6


def test_conservative_steps_never_increase_the_objective():
    """MMA's inner loop only accepts a candidate whose convex approximation dominates f there, so the sequence of
    outer iterates is monotone; the restatement returns the best point seen."""
    rng = np.random.default_rng(5)
    c = rng.uniform(0.5, 2.0, 5)
    f = lambda x: (float(np.sum(c * np.cosh(x - 1)) + 0.1 * np.sum(x ** 4)), c * np.sinh(x - 1) + 0.4 * x ** 3)
    x, fmin, nev, nouter, tr = mma(f, rng.normal(size=5) * 2)
    best = np.minimum.accumulate([v for _, v in tr])
    assert fmin == best[-1]
    ref = optimize.minimize(f, x, jac=True, method="BFGS", options=dict(gtol=1e-10))
    assert np.max(np.abs(x - ref.x)) <= 5e-3


def test_stop_rules_differ_only_in_when_they_stop():
    """NLopt <= 2.6 tests the x-tolerance per coordinate, >= 2.7 on L1 norms: same iterates, the 2.7 rule stops
    no later."""
    rng = np.random.default_rng(9)
    A = rng.normal(size=(5, 5))
    Q = A @ A.T + np.eye(5)
    b = rng.normal(size=5) * 3
    f = lambda x: (0.5 * x @ Q @ x - b @ x, Q @ x - b)
    _, _, n27, _, t27 = mma(f, np.ones(5), stop_rule=orc.STOP_NLOPT27)
    _, _, n26, _, t26 = mma(f, np.ones(5), stop_rule=orc.STOP_NLOPT26)
    m = min(n27, n26)
    assert all(np.array_equal(a[0], b_[0]) for a, b_ in zip(t27[:m], t26[:m]))
    assert n27 <= n26

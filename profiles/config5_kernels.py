import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, mmsig
D5 = 100_000
c5 = mmsig.synth.generate(D5, [7, 7], [96, 32])
rng = np.random.Generator(np.random.Philox(key=7))
g0s = rng.integers(1, 101, size=(8, 7 * 96 + 7 * 32)).astype(float)
for rep in range(2):
    m = mmsig.MMCTM([7, 7], [0.1, 0.1], c5, V=[96, 32], gamma0=g0s[0], profile=True)
    m.h.kernel_times(reset=True)
    t = time.perf_counter()
    elbo, ll, nit, best = m.fit_restarts(g0s, maxiter=30, tol=1e-4)
    dt = time.perf_counter() - t
    kt = m.h.kernel_times(reset=True)
    print("seconds", round(dt, 3), {k: (round(v[0], 1), v[1]) for k, v in kt.items() if v[0] > 1})
    m.close()

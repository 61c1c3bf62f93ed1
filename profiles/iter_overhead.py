#!/usr/bin/env python
"""Wall clock of mmsig_mmctm_iterate against the sum of its kernels at a small corpus (config 5 shape)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mmsig
D = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
c5 = mmsig.synth.generate(D, [7, 7], [96, 32])
g0 = mmsig.synth.init_gamma([7, 7], [96, 32])
for prof in (True, False):
    m = mmsig.MMCTM([7, 7], [0.1, 0.1], c5, V=[96, 32], gamma0=g0, profile=prof)
    for _ in range(5):
        m.iterate()
    m.h.kernel_times(reset=True)
    t = time.perf_counter()
    n = 100
    for _ in range(n):
        m.iterate()
    wall = (time.perf_counter() - t) / n * 1e3
    kt = m.h.kernel_times(reset=True)
    print("profile", prof, "wall ms/iter %.3f" % wall, "kernels ms/iter %.3f" % (sum(v[0] for v in kt.values()) / n),
          {k: round(v[0] / n, 3) for k, v in kt.items() if v[0] / n > 0.02})
    t = time.perf_counter()
    e = m.calculate_elbo()[0]
    print("  elbo call ms %.2f" % ((time.perf_counter() - t) * 1e3))
    t = time.perf_counter()
    m.set_state(g0)
    print("  set_state call ms %.2f" % ((time.perf_counter() - t) * 1e3))
    m.close()

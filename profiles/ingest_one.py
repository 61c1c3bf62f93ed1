#!/usr/bin/env python
"""One format_counts call per layout on the SNV-shaped modality (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mmsig import capi
from mmsig.counts import format_counts_device
D = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(1)
a = rng.poisson(30.0, (96, D)).astype(np.int32)
h = capi.Handle()
for layout in (0, 1):
    x = a if layout == 0 else np.ascontiguousarray(a.T)
    for _ in range(2):
        format_counts_device(x, layout=layout, handle=h)
h.close()

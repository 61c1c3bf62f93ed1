#!/usr/bin/env python
"""Count ingest on the device (format_counts_*, reference src/utils.jl:1-36): per-kernel time and
algorithmic HBM GB/s of k_dense_count / k_dense_fill for the three BASELINE modalities.
Usage (under gpurun): python profiles/bench_ingest.py [D]   -> one JSON object on stdout"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import mmsig  # noqa: E402
from mmsig import capi  # noqa: E402
from mmsig.counts import format_counts_device, make_count_csr  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
out = {"D": D, "hbm_peak_gbs": peak, "cases": []}
rng = np.random.default_rng(1)
for V, rate in [(96, 3500 / 96.0), (32, 85 / 32.0), (83, 300 / 83.0)]:
    base = rng.poisson(rate * rng.random((V, 1)) * 2 * (rng.random((V, D)) < 0.97), (V, D)).astype(np.int64)
    for dtype in (np.int32, np.int64):
        for layout in (0, 1):
            a = np.ascontiguousarray((base if layout == 0 else base.T).astype(dtype))
            h = capi.Handle(profile=True)
            format_counts_device(a, layout=layout, handle=h)          # warm-up (allocations, first launch)
            h.kernel_times(reset=True)
            t = time.perf_counter()
            n_rep = 3
            for _ in range(n_rep):
                r = format_counts_device(a, layout=layout, handle=h)
            wall = (time.perf_counter() - t) / n_rep
            kt = h.kernel_times(reset=True)
            h.close()
            nnz = int(r[0][-1])
            e = a.dtype.itemsize
            ms_c = kt["k_dense_count"][0] / kt["k_dense_count"][1]
            ms_f = kt["k_dense_fill"][0] / kt["k_dense_fill"][1]
            ms_s = kt["k_scan"][0] / kt["k_scan"][1]
            b_count = e * V * D + 8.0 * D
            b_fill = e * V * D + 8.0 * D + 8.0 * nnz
            out["cases"].append({
                "V": V, "dtype": a.dtype.name, "layout": "term-major" if layout == 0 else "sample-major", "nnz": nnz,
                "k_dense_count_ms": ms_c, "k_dense_count_GBs": b_count / ms_c / 1e6, "k_dense_count_frac": b_count / ms_c / 1e6 / peak,
                "k_dense_fill_ms": ms_f, "k_dense_fill_GBs": b_fill / ms_f / 1e6, "k_dense_fill_frac": b_fill / ms_f / 1e6 / peak,
                "k_scan_ms": ms_s, "end_to_end_ms_host_buffers": wall * 1e3,
                "host_numpy_make_count_csr_ms": None})
    t = time.perf_counter()
    ref = make_count_csr(base)
    out["cases"][-1]["host_numpy_make_count_csr_ms"] = (time.perf_counter() - t) * 1e3
    assert all(np.array_equal(x, y) for x, y in zip(r, ref))
print(json.dumps(out))

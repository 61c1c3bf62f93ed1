#!/bin/bash
# round 2, call 18: dense tile kernels with 2 / 4 samples per trip and no branch on the count (more independent chains per
# thread).  The -m gpu suite (bit-exactness of the MMCTM, 1e-12 of the LDA), then configs 4 and 2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -5 | tee gpurun_out/r02m_tests.log
for c in 4 2; do
  timeout 300 python bench.py --config $c --no-cpu --no-pageable --e2e-steps 2 > gpurun_out/r02m_bench_c$c.json 2> gpurun_out/r02m_bench_c$c.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r02m_bench_c$c.json"))
    print("config $c", "ms/it %.3f" % j["ms_per_step"], "value %.2f" % j["value"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()}, "e2e ms", j.get("e2e") and round(j["e2e"]["ms_per_step"], 2), "ll", j.get("ll"))
    if j.get("fp32_mode"):
        m = j["fp32_mode"]
        print("   fp32_mode: ms/it %.3f" % m["ms_per_step"], {k: round(x["ms_per_step"], 3) for k, x in m["kernels"].items()})
except Exception as e:
    print("config $c failed", e)
PY
done 2>&1 | tee gpurun_out/r02m_ab.log

#!/bin/bash
# round 2, call 16: longest-first sample order in the lean solver (counting sort by the previous iteration's evaluation
# counts) and the side-stream M-step half inside fit's sync-free loop.  The -m gpu suite, then A/B of MMSIG_ORDER=0 against
# the default at D = 1e6 and at D = 125000 (the shard of one rank of an 8-GPU run, where kernel tails weigh most).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -5 | tee gpurun_out/r02k_tests.log
for D in 1000000 125000; do
  for v in on off; do
    if [ $v = off ]; then export MMSIG_ORDER=0; else unset MMSIG_ORDER; fi
    timeout 200 python bench.py --samples $D --steps 10 --warmup 3 --no-cpu --no-pageable --no-fast --e2e-steps 1 > gpurun_out/r02k_D${D}_$v.json 2> gpurun_out/r02k_D${D}_$v.err
    python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r02k_D${D}_$v.json"))
    print("D=$D order $v", "ms/it %.3f (with kernel timing %.3f)" % (j["ms_per_step"], j["ms_per_step_with_kernel_timing"]), {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()}, "launches", j["gpu_launches"], "ll", j["ll"])
except Exception as e:
    print("D=$D order $v failed", e)
PY
  done
done 2>&1 | tee gpurun_out/r02k_ab.log

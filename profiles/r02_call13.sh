#!/bin/bash
# round 2, call 13: dense count tiles staged by cp.async.bulk + mbarrier in the four tile kernels (csrc/tile_stage.cuh).
# The whole -m gpu suite with the new default, then A/B of MMSIG_TILES=csr (round-2 kernels so far) against dense on
# configs 4 and 2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -8 | tee gpurun_out/r02h_tests.log
for v in dense csr; do
  if [ $v = csr ]; then export MMSIG_TILES=csr; else unset MMSIG_TILES; fi
  for c in 4 2; do
    timeout 300 python bench.py --config $c --no-cpu --no-pageable --e2e-steps 2 > gpurun_out/r02h_bench_c${c}_$v.json 2> gpurun_out/r02h_bench_c${c}_$v.err
    python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r02h_bench_c${c}_$v.json"))
    print("config $c $v", "ms/it %.3f" % j["ms_per_step"], "value %.2f" % j["value"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()}, "e2e", j.get("e2e") and round(j["e2e"].get("ms_per_step"), 2), "ll", j.get("ll"))
except Exception as e:
    print("config $c $v failed", e)
PY
  done
done 2>&1 | tee gpurun_out/r02h_ab.log
tail -3 gpurun_out/r02h_bench_c4_dense.err

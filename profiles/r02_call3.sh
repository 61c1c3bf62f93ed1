#!/bin/bash
# round 2, call 3: parity of the lean solver layouts + group path, A/B against the default, ncu of the lean kernels.
# Every step runs under its own timeout (a kernel that does not terminate must not hold the box).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mmctm.py -q -m gpu -x --timeout 120 -k "layouts or fit_host_rejects or literal" 2>&1 | tail -15 | tee gpurun_out/r02_call3_tests.log
timeout 300 python -m pytest tests/test_gpu_group.py tests/test_gpu_scale.py -q -m gpu -x --timeout 200 2>&1 | tail -15 | tee gpurun_out/r02_call3_tests2.log
D=1000000
for v in default lean8 lean4; do
  if [ "$v" = default ]; then unset MMSIG_SOLVE; else export MMSIG_SOLVE=$v; fi
  timeout 120 python bench.py --samples $D --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ab3_${v}.json 2> gpurun_out/ab3_${v}.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/ab3_${v}.json"))
    print("${v}", "ms/it %.3f" % j["ms_per_step"], {k: round(x["ms_per_step"], 3) for k, x in j["kernels"].items()}, "ll", j.get("ll"))
except Exception as e:
    print("${v} failed", e)
PY
done 2>&1 | tee gpurun_out/r02_call3_ab.log
export MMSIG_SOLVE=lean8
CMD="timeout 120 python bench.py --samples 400000 --steps 2 --warmup 1 --no-cpu --e2e-steps 1"
$CMD > gpurun_out/plain_r02a.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_solve_lean' -s 4 -c 2 -f -o gpurun_out/prof_r02a $CMD > gpurun_out/ncu_full_r02a.log 2>&1
tail -3 gpurun_out/ncu_full_r02a.log

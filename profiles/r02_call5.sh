#!/bin/bash
# round 2, call 5: lean solver with the staged sample pipeline (cp.async), FULL specialisation; A/B of blocks per SM and
# mat-vec unrolling; the small-model shapes (config 3 CTM K=10, config 5 shape) under lean8 / lean4 / pack; ncu.
mkdir -p gpurun_out
export MMSIG_SOLVE=lean8
timeout 600 python -m pytest tests/test_gpu_mmctm.py tests/test_gpu_group.py tests/test_gpu_heldout.py tests/test_gpu_immctm.py -q -m gpu -x --timeout 300 2>&1 | tail -6 | tee gpurun_out/r02_call5_tests.log
D=1000000
for v in default libmmsig_mb5 libmmsig_u6 libmmsig_u12; do
  if [ "$v" = default ]; then unset MMSIG_LIB; else export MMSIG_LIB=$PWD/multimodalmusig.jl_b200/$v.so; fi
  timeout 150 python bench.py --samples $D --steps 5 --warmup 3 --no-cpu --e2e-steps 1 --no-pageable > gpurun_out/ab5_${v}.json 2> gpurun_out/ab5_${v}.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/ab5_${v}.json"))
    print("${v}", "ms/it %.3f (profiled %.3f)" % (j["ms_per_step"], j["ms_per_step_with_kernel_timing"]), {k: round(x["ms_per_step"], 3) for k, x in j["kernels"].items()}, "ll", j.get("ll"), "e2e %.2f" % j["e2e"]["ms_per_step"], "fp64", j["roofline_fp64"] and round(j["roofline_fp64"]["frac"], 3))
except Exception as e:
    print("${v} failed", e)
PY
done 2>&1 | tee gpurun_out/r02_call5_ab.log
unset MMSIG_LIB
for v in warp lean8 lean4; do
  export MMSIG_SOLVE=$v
  timeout 150 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 --no-pageable > gpurun_out/ab5_c3_${v}.json 2> gpurun_out/ab5_c3_${v}.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/ab5_c3_${v}.json"))
    print("config3 ${v}", "ms/it %.3f" % j["ms_per_step"], {k: round(x["ms_per_step"], 3) for k, x in j["kernels"].items()})
except Exception as e:
    print("config3 ${v} failed", e)
PY
done 2>&1 | tee -a gpurun_out/r02_call5_ab.log
export MMSIG_SOLVE=lean8
CMD="timeout 150 python bench.py --samples 400000 --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --no-pageable"
$CMD > gpurun_out/plain_r02c.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_solve_lean' -s 4 -c 2 -f -o gpurun_out/prof_r02c $CMD > gpurun_out/ncu_full_r02c.log 2>&1
tail -2 gpurun_out/ncu_full_r02c.log

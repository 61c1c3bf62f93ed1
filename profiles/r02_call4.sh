#!/bin/bash
# round 2, call 4: full GPU suite; A/B of the lean solver at 3 / 4 / 5 blocks per SM (per-sample context in shared
# memory, 160 / 126 / 96 registers); ncu of the 4-block build.  Every step under its own timeout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -15 | tee gpurun_out/r02_call4_tests.log
D=1000000
export MMSIG_SOLVE=lean8
for v in default libmmsig_mb4 libmmsig_mb5; do
  if [ "$v" = default ]; then unset MMSIG_LIB; else export MMSIG_LIB=$PWD/multimodalmusig.jl_b200/$v.so; fi
  timeout 150 python bench.py --samples $D --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ab4_${v}.json 2> gpurun_out/ab4_${v}.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/ab4_${v}.json"))
    print("${v}", "ms/it %.3f" % j["ms_per_step"], {k: round(x["ms_per_step"], 3) for k, x in j["kernels"].items()}, "ll", j.get("ll"), "e2e %.2f" % j["e2e"]["ms_per_step"])
except Exception as e:
    print("${v} failed", e)
PY
done 2>&1 | tee gpurun_out/r02_call4_ab.log
export MMSIG_LIB=$PWD/multimodalmusig.jl_b200/libmmsig_mb4.so
CMD="timeout 150 python bench.py --samples 400000 --steps 2 --warmup 1 --no-cpu --e2e-steps 1"
$CMD > gpurun_out/plain_r02b.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_solve_lean' -s 4 -c 2 -f -o gpurun_out/prof_r02b $CMD > gpurun_out/ncu_full_r02b.log 2>&1
tail -2 gpurun_out/ncu_full_r02b.log

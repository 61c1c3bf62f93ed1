#!/bin/bash
# Round-2 A/B (under gpurun): the split-phase solver (csrc/mmctm_split.cuh, MMSIG_SOLVE=split) against the default
# k_solve on the bench workload, after its bit-exactness test.
#   bash profiles/ab_solve_split.sh [samples]
D=${1:-1000000}
mkdir -p gpurun_out
# the experimental kernels are not in the default build: make EXP=1 writes libmmsig_exp.so next to libmmsig.so
make -C multimodalmusig.jl_b200/csrc -s EXP=1 || exit 1
export MMSIG_LIB=$PWD/multimodalmusig.jl_b200/libmmsig_exp.so
MMSIG_EXPERIMENTAL=1 python -m pytest tests/test_gpu_mmctm.py -q -m gpu -k "split_phase or multi_sample" 2>&1 | tail -5 | tee gpurun_out/ab_split_parity.log
for v in default multi split split16; do
  if [ "$v" = default ]; then unset MMSIG_SOLVE; else export MMSIG_SOLVE=$v; fi
  python bench.py --samples $D --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ab_solve_${v}.json 2> gpurun_out/ab_solve_${v}.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/ab_solve_${v}.json"))
    print("${v}", "ms/it %.3f" % j["ms_per_step"], {k: round(x["ms_per_step"], 3) for k, x in j["kernels"].items()}, "ll", j.get("ll"))
except Exception as e:
    print("${v} failed", e)
PY
done

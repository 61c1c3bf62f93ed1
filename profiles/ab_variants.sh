#!/bin/bash
# A/B timing of libmmsig build variants on the bench workload (under gpurun).
# Usage: bash profiles/ab_variants.sh <tag> <samples> <variant.so> [<variant.so> ...]   ("default" = in-tree libmmsig.so)
TAG=$1; D=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  name=$(basename "$v" .so)
  if [ "$v" = default ]; then unset MMSIG_LIB; else export MMSIG_LIB=$PWD/$v; fi
  python bench.py --samples $D --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ab_${TAG}_${name}.json 2> gpurun_out/ab_${TAG}_${name}.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/ab_${TAG}_${name}.json"))
    print("${name}", "ms/it %.3f" % j["ms_per_step"], {k: round(v["ms_per_step"], 3) for k, v in j["kernels"].items()}, "ll", j["ll"])
except Exception as e:
    print("${name} failed", e)
PY
done

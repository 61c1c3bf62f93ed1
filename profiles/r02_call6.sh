#!/bin/bash
# round 2, call 6: lean solver as the default, sync-free early iterations, packed e2e; the bench lines of configs 4, 3, 2, 5, 1
mkdir -p gpurun_out
rm -f multimodalmusig.jl_b200/libmmsig_mb5.so multimodalmusig.jl_b200/libmmsig_u6.so multimodalmusig.jl_b200/libmmsig_u12.so
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -6 | tee gpurun_out/r02_call6_tests.log
timeout 400 python bench.py > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; tail -c 1500 gpurun_out/r02_bench_c4.json
for c in 3 2 1; do
  timeout 300 python bench.py --config $c > gpurun_out/r02_bench_c$c.json 2> gpurun_out/r02_bench_c$c.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r02_bench_c$c.json"))
    print("config $c", "ms/it %.3f" % j["ms_per_step"], "value %.2f" % j["value"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()}, "roof", j.get("roofline") and round(j["roofline"]["frac"], 4), "cpu", j.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("config $c failed", e)
PY
done 2>&1 | tee gpurun_out/r02_call6_configs.log
for v in lean8 lean4 warp; do
  export MMSIG_SOLVE=$v
  timeout 300 python bench.py --config 5 --gpus 1 --steps 1 > gpurun_out/r02_bench_c5_$v.json 2> gpurun_out/r02_bench_c5_$v.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r02_bench_c5_$v.json"))
    print("config 5 $v", "s/batch %.3f" % (j["ms_per_step"] / 1e3), "it/s %.1f" % j["value"], "best", j["best_restart"], j["best_elbo"])
except Exception as e:
    print("config 5 $v failed", e)
PY
done 2>&1 | tee -a gpurun_out/r02_call6_configs.log

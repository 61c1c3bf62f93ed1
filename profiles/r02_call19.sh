#!/bin/bash
# round 2, call 19: the stopping rule on the device / batched iterations beyond the tenth.  The -m gpu suite (golden fits,
# LL histories and stop positions bit-exact against the oracle), config 1 (100-iteration fit of 560 samples: all launch
# latency and host round trips) and config 4 with 20 steps (10 sync-free + 10 batched).
mkdir -p gpurun_out
export MMSIG_TEST_DEVICE_RULE=1          # run the device-rule case of test_fit_stops_exactly_where_the_reference_rule_fires
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -5 | tee gpurun_out/r02n_tests.log
for v in 1 0; do
export MMSIG_DEVICE_RULE=$v
timeout 200 python bench.py --config 1 > gpurun_out/r02n_bench_c1_rule$v.json 2> gpurun_out/r02n_bench_c1_rule$v.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-pageable --no-fast --e2e-steps 1 > gpurun_out/r02n_bench_c4_20_rule$v.json 2> gpurun_out/r02n_bench_c4_20_rule$v.err
done
MMSIG_DEVICE_RULE=1 timeout 600 python -m pytest tests/test_gpu_mmctm.py tests/test_golden_fits.py tests/test_gpu_immctm.py tests/test_gpu_group.py tests/test_gpu_heldout.py -q -m gpu -x --timeout 300 2>&1 | tail -3 | tee gpurun_out/r02n_tests_rule1.log
python - <<'PY'
import json
for f in ("c1_rule1", "c1_rule0", "c4_20_rule1", "c4_20_rule0"):
    try:
        j = json.load(open("gpurun_out/r02n_bench_%s.json" % f))
        print(f, "ms/it %.4f" % j["ms_per_step"], "value %.2f" % j["value"], "steps", j["steps"], "launches", j.get("gpu_launches"), "ll", j.get("ll"), "converged", j.get("converged"))
    except Exception as e:
        print(f, "failed", e)
PY

#!/bin/bash
# round 2, call 14: the optional FP32 mode of the tile passes (csrc/tile_f32.cuh): its tests, the default bench line
# (FP64 headline + the fp32_mode field), config 2 in both precisions; the suite once more (k_pack_rows writes dense rows).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fp32.py -q -m gpu -x -s --timeout 300 2>&1 | grep -E "fp32|passed|failed|Error|assert" | tail -30 | tee gpurun_out/r02i_fp32_tests.log
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -5 | tee gpurun_out/r02i_tests.log
timeout 300 python bench.py --no-cpu --no-pageable --e2e-steps 2 > gpurun_out/r02i_bench_c4.json 2> gpurun_out/r02i_bench_c4.err
timeout 300 python bench.py --config 2 --no-cpu > gpurun_out/r02i_bench_c2.json 2> gpurun_out/r02i_bench_c2.err
timeout 300 python bench.py --config 2 --no-cpu --precision fp32 > gpurun_out/r02i_bench_c2_fp32.json 2> gpurun_out/r02i_bench_c2_fp32.err
python - <<'PY'
import json
for f in ("c4", "c2", "c2_fp32"):
    try:
        j = json.load(open("gpurun_out/r02i_bench_%s.json" % f))
        print(f, "ms/it %.3f" % j["ms_per_step"], "value %.2f" % j["value"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()}, "e2e ms", j.get("e2e") and round(j["e2e"]["ms_per_step"], 2), "ll", j.get("ll"))
        if j.get("fp32_mode"):
            m = j["fp32_mode"]
            print("   fp32_mode: ms/it %.3f" % m["ms_per_step"], {k: round(x["ms_per_step"], 3) for k, x in m["kernels"].items()}, "ll", m["ll"], "rel diff", m["ll_rel_diff_to_fp64"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r02i_bench_c4.err

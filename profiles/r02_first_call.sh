#!/bin/bash
# One gpurun call that opens the next round (B200, one GPU, ~4 min of box time):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/r02_first_call.sh'
# 1. the GPU suite on the default library, 2. the FP64 pipe's rate (the E-step's second roof),
# 3. parity + A/B of the experimental split-phase solver, 4. the default bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests.log
timeout 120 bash profiles/micro/run_fp64_peak.sh 2>&1 | tail -20
timeout 400 bash profiles/ab_solve_split.sh 1000000 2>&1 | tail -12 | tee gpurun_out/r02_ab_solve_split.log
timeout 300 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 600 gpurun_out/r02_bench_default.json

#!/bin/bash
# FP64 pipe rate of the box's GPU (under gpurun): bash profiles/micro/run_fp64_peak.sh
set -e
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/fp64_peak profiles/micro/fp64_peak.cu
./gpurun_out/fp64_peak | tee gpurun_out/fp64_peak.jsonl

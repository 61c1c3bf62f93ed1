#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one full
# capture of the hot kernels of one iteration.  Usage (under gpurun): bash profiles/run_ncu.sh <tag> [samples]
set -u
TAG=${1:-r01}
D=${2:-200000}
CMD="python bench.py --samples $D --steps 2 --warmup 1 --no-cpu --e2e-steps 1"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_theta_tile|k_loglik_tile|k_moments' -s 8 -c 8 \
    -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/
tail -3 gpurun_out/ncu_full_$TAG.log

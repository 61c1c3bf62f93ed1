#!/bin/bash
# e2e (mmsig_mmctm_fit_host, maxiter = 1) against the number of pipeline chunks / their growth ratio
for cfg in "7 1.0" "12 1.0" "20 1.0" "12 1.3" "20 1.2"; do
  set -- $cfg
  MMSIG_PIPE_CHUNKS=$1 MMSIG_PIPE_RATIO=$2 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 4 > gpurun_out/chunks_$1_$2.json 2> gpurun_out/chunks_$1_$2.err
  python -c "
import json; j=json.load(open('gpurun_out/chunks_$1_$2.json')); print('chunks $1 ratio $2', 'value ms', round(j['ms_per_step'],2), 'e2e ms', round(j['e2e']['ms_per_step'],2))"
done

#!/bin/bash
# round 2, call 21 (2 GPUs): the batched fit loop (stopping rule on the device, default on) under NCCL: the two-rank tests
# with the new whole-fit mode, and the torchrun bench at N = 2 with 20 steps (10 sync-free + 10 batched).
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -q -m gpu -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02q_tests.log
export BENCH_TRACE=100
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-pageable > gpurun_out/r02q_torchrun_n2_20.json 2> gpurun_out/r02q_torchrun_n2_20.err
echo "torchrun rc=$?"
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/r02q_torchrun_n2_20.json"))
    print("torchrun n2 steps 20: ms/it %.3f value %.2f" % (j["ms_per_step"], j["value"]), "e2e %.2f ms" % j["e2e"]["ms_per_step"], "ll", j["ll"])
except Exception as e:
    print("failed", e)
PY

#!/bin/bash
# round 2, call 23 (8 GPUs): the torchrun arm at N = 8 once more with the side-stream M-step half and the batched fit loop
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --no-cpu --no-pageable --e2e-steps 2 > gpurun_out/r02t_torchrun_n8.json 2> gpurun_out/r02t_torchrun_n8.err
echo "rc=$?"
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/r02t_torchrun_n8.json"))
    print("torchrun n8: ms/it %.3f (with kernel timing %.3f) value %.2f" % (j["ms_per_step"], j["ms_per_step_with_kernel_timing"], j["value"]), "e2e %.2f ms (%.1f it/s)" % (j["e2e"]["ms_per_step"], j["e2e"]["value"]), "ll", j["ll"])
except Exception as e:
    print("failed", e)
PY

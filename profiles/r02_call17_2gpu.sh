#!/bin/bash
# round 2, call 17 (2 GPUs): the side-stream M-step half (gather 2 on the side stream) under NCCL: the two-rank parity
# tests, the new tile-format test, and the torchrun bench at N = 2 (12.63 ms per iteration before).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py tests/test_gpu_mmctm.py -q -m gpu -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02l_tests.log
export BENCH_TRACE=100
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu --no-pageable > gpurun_out/r02l_torchrun_n2.json 2> gpurun_out/r02l_torchrun_n2.err
echo "torchrun rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --no-cpu --no-pageable --samples 250000 > gpurun_out/r02l_torchrun_n2_250k.json 2> gpurun_out/r02l_torchrun_n2_250k.err
python - <<'PY'
import json
for f in ("torchrun_n2", "torchrun_n2_250k"):
    try:
        j = json.load(open("gpurun_out/r02l_%s.json" % f))
        print(f, "ms/it %.3f (with kernel timing %.3f) value %.2f" % (j["ms_per_step"], j["ms_per_step_with_kernel_timing"], j["value"]), "e2e %.2f ms" % j["e2e"]["ms_per_step"], "ll", j["ll"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY

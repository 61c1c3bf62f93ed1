#!/bin/bash
# round 2, call 12 (2 GPUs): the torchrun arm of the bench at N = 2 did not finish within 300 s in call 11 (no traceback, both
# ranks alive).  Same command at D = 200000 with stage marks and a Python stack dump every 60 s (BENCH_TRACE), then at full size.
mkdir -p gpurun_out
export BENCH_TRACE=60 NCCL_DEBUG=WARN
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu --no-pageable --samples 200000 --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/r02g_torchrun_n2_small.json 2> gpurun_out/r02g_torchrun_n2_small.err
echo "small: rc=$?"; grep -E "^\[bench|File|Thread|Current" gpurun_out/r02g_torchrun_n2_small.err | tail -60
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --no-cpu --no-pageable > gpurun_out/r02g_torchrun_n2.json 2> gpurun_out/r02g_torchrun_n2.err
echo "full: rc=$?"; grep -E "^\[bench" gpurun_out/r02g_torchrun_n2.err | tail -30
python - <<'PY'
import json
for f in ("torchrun_n2_small", "torchrun_n2"):
    try:
        j = json.load(open("gpurun_out/r02g_%s.json" % f))
        print(f, "ms/it %.3f value %.2f" % (j["ms_per_step"], j["value"]), "e2e %.2f ms" % j["e2e"]["ms_per_step"], "ll", j["ll"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY

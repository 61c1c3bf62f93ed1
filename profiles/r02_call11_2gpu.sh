#!/bin/bash
# round 2, call 11 (2 GPUs): the whole `-m gpu` suite on a two-GPU lease (the two-GPU parity tests included: NCCL one process
# per GPU; single-process group over two devices), with the register-resident LU of k_mstep2 and the four-loads-in-flight
# k_combine; the bench at N = 2 both ways
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -6 | tee gpurun_out/r02_call11_tests.log
timeout 300 python bench.py --gpus 2 --no-cpu --no-pageable > gpurun_out/r02f_bench_group_n2.json 2> gpurun_out/r02f_bench_group_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu --no-pageable > gpurun_out/r02f_bench_torchrun_n2.json 2> gpurun_out/r02f_bench_torchrun_n2.err
python - <<'PY'
import json
for f in ("group_n2", "torchrun_n2"):
    try:
        j = json.load(open("gpurun_out/r02f_bench_%s.json" % f))
        print(f, "ms/it %.3f value %.2f" % (j["ms_per_step"], j["value"]), "e2e %.2f ms" % j["e2e"]["ms_per_step"], "ll", j["ll"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY
tail -5 gpurun_out/r02f_bench_group_n2.err gpurun_out/r02f_bench_torchrun_n2.err

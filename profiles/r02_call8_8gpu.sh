#!/bin/bash
# round 2, call 8 (8 GPUs): the bench at N = 4 and 8 under torchrun (one process per GPU, NCCL) and at N = 8 as ONE process
# (group API, peer-memory exchange); config 5 (64 restarts dealt over 8 GPUs); group parity test over 4 distinct devices
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_group.py -q -m gpu -x --timeout 200 -k "bit_identical or restarts" 2>&1 | tail -4 | tee gpurun_out/r02_call8_tests.log
for n in 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu --no-pageable > gpurun_out/r02_bench_torchrun_n$n.json 2> gpurun_out/r02_bench_torchrun_n$n.err
done
timeout 300 python bench.py --gpus 8 --no-cpu --no-pageable > gpurun_out/r02_bench_group_n8.json 2> gpurun_out/r02_bench_group_n8.err
timeout 300 python bench.py --config 5 --gpus 8 --steps 1 > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err
python - <<'PY'
import json
for f in ("torchrun_n4", "torchrun_n8", "group_n8"):
    try:
        j = json.load(open("gpurun_out/r02_bench_%s.json" % f))
        print(f, "ms/it %.3f value %.2f" % (j["ms_per_step"], j["value"]), "e2e %.2f ms (%.1f it/s)" % (j["e2e"]["ms_per_step"], j["e2e"]["value"]), "link %.1f" % j["e2e"]["h2d_link_gbs_measured"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()})
    except Exception as e:
        print(f, "failed", e)
try:
    j = json.load(open("gpurun_out/r02_bench_c5_n8.json"))
    print("config 5 on 8 GPUs: %.3f s per batch of %d restarts x %d iterations, %.1f it/s, best %d elbo %.3f" % (j["ms_per_step"] / 1e3, j["restarts"], j["iterations_per_restart"], j["value"], j["best_restart"], j["best_elbo"]))
except Exception as e:
    print("config 5 failed", e)
PY

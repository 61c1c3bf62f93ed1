#!/bin/bash
# round 2, call 15 (8 GPUs): the bench at N = 8 and 4 under torchrun (one process per GPU, NCCL: the driver's SCALE path), at
# N = 8 as ONE process (mmsig_group_*, peer-memory exchange), config 5 (64 restarts dealt over 8 GPUs), and the group
# parity tests over distinct devices.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
export BENCH_TRACE=100
for n in 8 4; do
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu --no-pageable > gpurun_out/r02j_torchrun_n$n.json 2> gpurun_out/r02j_torchrun_n$n.err
  echo "torchrun n=$n rc=$?"
done
timeout 170 python bench.py --gpus 8 --no-cpu --no-pageable > gpurun_out/r02j_group_n8.json 2> gpurun_out/r02j_group_n8.err; echo "group n=8 rc=$?"
timeout 170 python bench.py --config 5 --gpus 8 --steps 1 > gpurun_out/r02j_c5_n8.json 2> gpurun_out/r02j_c5_n8.err; echo "config 5 rc=$?"
timeout 200 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -q -m gpu -x --timeout 150 2>&1 | tail -4 | tee gpurun_out/r02j_tests.log
python - <<'PY'
import json
for f in ("torchrun_n8", "torchrun_n4", "group_n8"):
    try:
        j = json.load(open("gpurun_out/r02j_%s.json" % f))
        print(f, "ms/it %.3f value %.2f" % (j["ms_per_step"], j["value"]), "e2e %.2f ms (%.1f it/s)" % (j["e2e"]["ms_per_step"], j["e2e"]["value"]), "link %.1f" % j["e2e"]["h2d_link_gbs_measured"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()}, "ll", j["ll"])
    except Exception as e:
        print(f, "failed", e)
try:
    j = json.load(open("gpurun_out/r02j_c5_n8.json"))
    print("config 5 on 8 GPUs: %.3f s per batch of %d restarts x %d iterations, %.1f it/s, best %d elbo %.3f" % (j["ms_per_step"] / 1e3, j["restarts"], j["iterations_per_restart"], j["value"], j["best_restart"], j["best_elbo"]))
except Exception as e:
    print("config 5 failed", e)
PY
grep -h "^\[bench" gpurun_out/r02j_torchrun_n8.err | tail -4

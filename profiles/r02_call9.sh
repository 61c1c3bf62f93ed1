#!/bin/bash
# round 2, call 9 (session 2): the round's evidence in one call -- the whole `-m gpu` suite, the plain bench lines of
# configs 4 (default), 2 and 3, the ncu launch list of the default bench command, one `--set full` capture of one
# iteration's hot MMCTM kernels at D = 1e6 and one of the LDA kernels (config 2 shape).
mkdir -p gpurun_out
T=r02d
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -6 | tee gpurun_out/${T}_tests.log
timeout 500 python bench.py > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; tail -c 600 gpurun_out/${T}_bench_c4.json
for c in 2 3; do
  timeout 300 python bench.py --config $c > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err
done
python - <<'PY'
import json
for c in (4, 2, 3):
    try:
        j = json.load(open("gpurun_out/r02d_bench_c%d.json" % c))
        print("config", c, "ms/it %.3f" % j["ms_per_step"], "value %.2f" % j["value"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()},
              "roof", j.get("roofline") and round(j["roofline"]["frac"], 4), "fp64", j.get("roofline_fp64") and round(j["roofline_fp64"]["frac"], 3),
              "e2e", j.get("e2e") and j["e2e"].get("value"), "cpu", j.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print("config", c, "failed", e)
PY
CMD="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-pageable"
$CMD > gpurun_out/plain_$T.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$T.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launches_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_theta_tile|k_loglik_tile|k_moments' -s 16 -c 9 \
    -f -o gpurun_out/prof_$T $CMD > gpurun_out/ncu_full_$T.log 2>&1
tail -2 gpurun_out/ncu_full_$T.log
cat > /tmp/lda_prof.py <<PY
import sys; sys.path.insert(0, "$PWD")
import mmsig
csr = mmsig.synth.generate(1000000, [20], [96])[0]
m = mmsig.LDA(20, 0.1, 0.1, csr, V=96, lambda0=mmsig.synth.init_lda_lambda(20, 96))
for _ in range(4): print(m.iterate())
PY
python /tmp/lda_prof.py > gpurun_out/plain_${T}_lda.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lda_' -s 6 -c 4 -f -o gpurun_out/prof_${T}_lda python /tmp/lda_prof.py > gpurun_out/ncu_${T}_lda.log 2>&1
tail -2 gpurun_out/ncu_${T}_lda.log
ls -la gpurun_out/

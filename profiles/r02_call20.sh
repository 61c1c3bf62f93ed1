#!/bin/bash
# round 2, call 20: the round's final evidence.  Plain bench lines (configs 4 with the CPU arm, 2, 3), then the ncu launch
# list of the default bench command and one `--set full` capture each of: one iteration's hot MMCTM kernels at D = 1e6 (dense
# bulk-staged theta tiles of the three modalities, k_solve_lean nu / lambda, k_moments, k_loglik_tile), the LDA's two FP64
# kernels and its two FP32 kernels.  Raw / source pages become CSV on the box (the reports exceed the 64 MiB pull limit).
mkdir -p gpurun_out
T=r02p
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -5 | tee gpurun_out/${T}_tests.log
timeout 400 python bench.py > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err
for c in 2 3; do timeout 200 python bench.py --config $c --no-cpu > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err; done
python - <<'PY'
import json
for c in (4, 2, 3):
    try:
        j = json.load(open("gpurun_out/r02p_bench_c%d.json" % c))
        print("config", c, "ms/it %.3f" % j["ms_per_step"], "value %.2f" % j["value"], {k: round(x["ms_per_step"], 3) for k, x in (j.get("kernels") or {}).items()},
              "roof", j.get("roofline") and round(j["roofline"]["frac"], 4), "fp64", j.get("roofline_fp64") and round(j["roofline_fp64"]["frac"], 3),
              "e2e", j.get("e2e") and round(j["e2e"].get("value"), 2), "cpu", j.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print("config", c, "failed", e)
PY
CMD="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-fast --e2e-steps 1 --no-pageable"
$CMD > gpurun_out/plain_$T.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$T.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launches_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_theta_tile|k_loglik_tile|k_moments' -s 27 -c 7 \
    -f -o gpurun_out/prof_$T $CMD > gpurun_out/ncu_full_$T.log 2>&1
tail -2 gpurun_out/ncu_full_$T.log
for prec in fp64 fp32; do
cat > /tmp/lda_prof_$prec.py <<PY
import sys; sys.path.insert(0, "$PWD")
import mmsig
csr = mmsig.synth.generate(1000000, [20], [96])[0]
m = mmsig.LDA(20, 0.1, 0.1, csr, V=96, lambda0=mmsig.synth.init_lda_lambda(20, 96), precision="$prec")
for _ in range(4): print(m.iterate())
PY
python /tmp/lda_prof_$prec.py > gpurun_out/plain_${T}_lda_$prec.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lda_estep|k_lda_ll' -s 4 -c 2 -f -o gpurun_out/prof_${T}_lda_$prec python /tmp/lda_prof_$prec.py > gpurun_out/ncu_${T}_lda_$prec.log 2>&1
tail -1 gpurun_out/ncu_${T}_lda_$prec.log
done
for r in prof_$T prof_${T}_lda_fp64 prof_${T}_lda_fp32; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2> /dev/null
  ncu -i gpurun_out/$r.ncu-rep --page source --csv > gpurun_out/${r}_source.csv 2> /dev/null
done
gzip -9 gpurun_out/*_source.csv
rm -f gpurun_out/prof_${T}_lda_fp64.ncu-rep gpurun_out/prof_${T}_lda_fp32.ncu-rep
if [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; then rm -f gpurun_out/prof_$T.ncu-rep; fi
du -sm gpurun_out; ls gpurun_out | head -40

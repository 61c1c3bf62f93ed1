#!/bin/bash
# round 2, call 10: the ncu evidence of call 9 again, sized to come back (gpurun pulls at most 64 MiB): launch list of the
# default bench command, one `--set full` capture of one iteration's hot MMCTM kernels at D = 1e6 (first modality's tile
# kernels, both solve phases, moments) and one of the two LDA kernels (config 2 shape).  Raw and source pages are turned
# into CSV on the box; the .ncu-rep files are dropped if the directory would exceed the limit.
mkdir -p gpurun_out
T=r02e
CMD="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-pageable"
$CMD > gpurun_out/plain_$T.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$T.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launches_$T.log 2>&1
# launches per iteration: theta x3, solve x2, combine, mstep1, moments, loglik x3, combine, mstep2 = 13; skip 4 iterations
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_theta_tile|k_loglik_tile|k_moments' -s 27 -c 7 \
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lda_estep|k_lda_ll_tile' -s 4 -c 2 -f -o gpurun_out/prof_${T}_lda python /tmp/lda_prof.py > gpurun_out/ncu_${T}_lda.log 2>&1
tail -2 gpurun_out/ncu_${T}_lda.log
for r in prof_$T prof_${T}_lda; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2> /dev/null
  ncu -i gpurun_out/$r.ncu-rep --page source --csv > gpurun_out/${r}_source.csv 2> /dev/null
done
gzip -9 gpurun_out/*_source.csv
du -sm gpurun_out
if [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; then rm -f gpurun_out/prof_${T}_lda.ncu-rep; fi
if [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; then rm -f gpurun_out/prof_$T.ncu-rep; fi
ls -la gpurun_out/

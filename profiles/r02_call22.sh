#!/bin/bash
# round 2, call 22: `--set full` of the LDA's E-pass kernels (call 20's skip count landed on the LL kernels only): the dense
# bulk-staged FP64 kernel and the FP32 kernel, config 2 shape.
mkdir -p gpurun_out
T=r02s
for prec in fp64 fp32; do
cat > /tmp/lda_prof_$prec.py <<PY
import sys; sys.path.insert(0, "$PWD")
import mmsig
csr = mmsig.synth.generate(1000000, [20], [96])[0]
m = mmsig.LDA(20, 0.1, 0.1, csr, V=96, lambda0=mmsig.synth.init_lda_lambda(20, 96), precision="$prec")
for _ in range(3): print(m.iterate())
PY
python /tmp/lda_prof_$prec.py > gpurun_out/plain_${T}_lda_$prec.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_lda_estep' -s 2 -c 1 -f -o gpurun_out/prof_${T}_lda_$prec python /tmp/lda_prof_$prec.py > gpurun_out/ncu_${T}_lda_$prec.log 2>&1
tail -1 gpurun_out/ncu_${T}_lda_$prec.log
ncu -i gpurun_out/prof_${T}_lda_$prec.ncu-rep --page raw --csv > gpurun_out/prof_${T}_lda_${prec}_raw.csv 2> /dev/null
rm -f gpurun_out/prof_${T}_lda_$prec.ncu-rep
done
ls gpurun_out

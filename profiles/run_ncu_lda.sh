#!/bin/bash
# one full capture of the LDA kernels (config 2 shape) -- usage under gpurun: bash profiles/run_ncu_lda.sh <tag> [D]
TAG=${1:-lda}
D=${2:-200000}
cat > /tmp/lda_prof.py <<PY
import sys; sys.path.insert(0, "$PWD")
import mmsig
csr = mmsig.synth.generate($D, [20], [96])[0]
m = mmsig.LDA(20, 0.1, 0.1, csr, V=96, lambda0=mmsig.synth.init_lda_lambda(20, 96))
for _ in range(4): print(m.iterate())
PY
python /tmp/lda_prof.py > gpurun_out/plain_$TAG.log 2>&1 || { tail gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_lda_estep' -s 2 -c 1 -f -o gpurun_out/prof_$TAG python /tmp/lda_prof.py > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log

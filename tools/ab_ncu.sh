#!/bin/bash
# ncu --set full of the curve kernel with the plain and the SASS-patched library (single-key verify, 2^20 tuples)
set -u
for v in "" _sp5; do
  rep=/tmp/ab${v}
  SB200_LIB=$PWD/schnorr_b200/libschnorr_b200${v}.so ncu --set full --clock-control none --kernel-name-base demangled -k regex:k_run --launch-skip 2 -c 2 -f -o $rep python tools/prof_op.py verify 20 1 > gpurun_out/ab_ncu${v}.log 2>&1
  python tools/ncu_summary.py $rep.ncu-rep > gpurun_out/ab_ncu${v}_summary.txt 2>&1
done
tail -3 gpurun_out/ab_ncu_sp5.log

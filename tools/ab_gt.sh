#!/bin/bash
set -u
out=gpurun_out/ab_gt.log; : > $out
for v in "" _gt _gtp "" _gtp; do
  echo "=== lib${v}" >> $out
  SB200_LIB=$PWD/schnorr_b200/libschnorr_b200${v}.so python tools/prof_op.py verify 22 3 >> $out 2>&1
done
SB200_LIB=$PWD/schnorr_b200/libschnorr_b200_gtp.so ncu --set full --clock-control none --kernel-name-base demangled -k regex:k_verify_ec_p --launch-skip 1 -c 1 -f -o /tmp/gt python tools/prof_op.py verify 20 1 > gpurun_out/ab_gt_ncu.log 2>&1
python tools/ncu_summary.py /tmp/gt.ncu-rep | grep -v fp64 > gpurun_out/ab_gt_ncu_summary.txt 2>&1
cat $out; head -40 gpurun_out/ab_gt_ncu_summary.txt

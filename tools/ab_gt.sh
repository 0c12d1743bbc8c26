#!/bin/bash
# A/B of the window-table placement for the curve kernel: local memory (default) vs thread-major global tables
# (-DSB_EC_GLOBAL_TABLES=1) with the lookups staged into shared memory by cp.async (-DSB_TABLE_STAGE=1); timings, verdict
# check, and ncu of the staged kernel.   usage: tools/ab_gt.sh <suffix-of-the-variant-library>
set -u
v=${1:-_gts}
out=gpurun_out/ab_gt.log; : > $out
for lib in "" $v "" $v; do
  echo "=== lib${lib}" >> $out
  SB200_LIB=$PWD/schnorr_b200/libschnorr_b200${lib}.so python tools/prof_op.py verify 22 3 >> $out 2>&1
done
echo "=== pytest lib${v} (single-key verify tests)" >> $out
SB200_LIB=$PWD/schnorr_b200/libschnorr_b200${v}.so timeout 900 python -m pytest tests -m gpu -q -k "verify_single or small_order or config1 or golden or half_size or sign_and_verify" >> $out 2>&1
SB200_LIB=$PWD/schnorr_b200/libschnorr_b200${v}.so ncu --set full --clock-control none --kernel-name-base demangled -k regex:k_curve_p --launch-skip 1 -c 1 -f -o /tmp/gt python tools/prof_op.py verify 20 1 > gpurun_out/ab_gt_ncu.log 2>&1
python tools/ncu_summary.py /tmp/gt.ncu-rep | grep -v fp64 > gpurun_out/ab_gt_ncu_summary.txt 2>&1
grep -E "===|verify n|passed|failed" $out; grep -E "gpu__time|dram|issue_active|long_scoreboard|short_scoreboard|lg_throttle|warps_active" gpurun_out/ab_gt_ncu_summary.txt

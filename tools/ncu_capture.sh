#!/bin/bash
# ncu --set full capture of one op's kernels on the GPU box, reduced to text there (the .ncu-rep files are 10-15 MB each and
# gpurun brings back at most 64 MiB): raw-page summary, per-opcode table of every captured kernel, DRAM traffic.
# usage: tools/ncu_capture.sh <op> <log2n> <kernel-regex> <traffic-key> <tag> [launch-skip]   (writes gpurun_out/ncu_<op>_<tag>_*.txt)
# (kernels longer than ~100 ms come back with NaN counters from `--set full`: capture single-key verify at 2^21, not 2^22)
set -u
op=$1; lg=$2; kre=$3; key=$4; tag=$5; skip=${6:-0}
rep=/tmp/${op}_${tag}.ncu-rep
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$kre --launch-skip $skip -c 2 -f -o ${rep%.ncu-rep} python tools/prof_op.py $op $lg 1 > gpurun_out/ncu_${op}_${tag}.log 2>&1
python tools/ncu_summary.py $rep > gpurun_out/ncu_${op}_${tag}_summary.txt 2>&1
for id in 1 2; do
  ncu -i $rep --page source --csv --kernel-id :::$id > /tmp/src_${op}_$id.csv 2>/dev/null
  if [ -s /tmp/src_${op}_$id.csv ]; then
    head -1 /tmp/src_${op}_$id.csv >> gpurun_out/ncu_${op}_${tag}_opmix.txt
    python tools/ncu_opmix.py /tmp/src_${op}_$id.csv >> gpurun_out/ncu_${op}_${tag}_opmix.txt 2>&1
    kn=$(head -1 /tmp/src_${op}_$id.csv | sed -E 's/.*(k_run|k_fixed_batch|k_curve_p)<\(int\)([0-9]+).*/\1ILi\2E/')
    echo "## kernel $id ($kn): executed warp instructions and stall samples per device function" >> gpurun_out/ncu_${op}_${tag}_byfunc.txt
    python tools/ncu_by_func.py /tmp/src_${op}_$id.csv ${SB200_LIB:-schnorr_b200/libschnorr_b200.so} "$kn" >> gpurun_out/ncu_${op}_${tag}_byfunc.txt 2>&1
  fi
done
python tools/ncu_traffic.py $key $((1 << lg)) $rep gpurun_out/ncu_traffic_${tag}.json > /dev/null 2>&1
tail -1 gpurun_out/ncu_${op}_${tag}.log

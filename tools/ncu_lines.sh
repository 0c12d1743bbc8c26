#!/bin/bash
# per-source-line executed instructions of the curve kernel (development aid)
set -u
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_run --launch-skip 3 -c 1 -f -o /tmp/ln python tools/prof_op.py verify 20 1 > gpurun_out/ncu_lines.log 2>&1
ncu -i /tmp/ln.ncu-rep --page source --print-source cuda --csv > gpurun_out/ncu_lines_cuda.csv 2>>gpurun_out/ncu_lines.log
ncu -i /tmp/ln.ncu-rep --page source --csv > /tmp/src_ln.csv 2>/dev/null
python tools/ncu_by_func.py /tmp/src_ln.csv schnorr_b200/libschnorr_b200.so k_runILi19E > gpurun_out/ncu_lines_byfunc.txt 2>&1
wc -l gpurun_out/ncu_lines_cuda.csv; head -c 1500 gpurun_out/ncu_lines_cuda.csv; cat gpurun_out/ncu_lines_byfunc.txt

#!/bin/bash
# A/B of the SASS pass (tools/sass_patch.py) on the GPU box: device-resident timings of every op with the plain and the patched
# libraries, then the whole GPU parity suite on the patched one.
set -u
out=gpurun_out/ab_sass.log; : > $out
for v in "" _sp5 _sp6 _sp5b "" _sp5 _sp5b; do
  echo "=== lib${v}" >> $out
  SB200_LIB=$PWD/schnorr_b200/libschnorr_b200${v}.so python tools/quick_time.py $((1<<22)) >> $out 2>&1
done
for v in _sp5b _sp5; do
  echo "=== pytest lib${v}" >> $out
  SB200_LIB=$PWD/schnorr_b200/libschnorr_b200${v}.so timeout 1500 python -m pytest tests -m gpu -x -q >> $out 2>&1
done
tail -5 $out

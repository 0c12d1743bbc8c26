#!/bin/bash
# A/B of library variants on the GPU box: tools/ab_libs.sh <log2n> <suffix> [<suffix> ...]   ("" = the default library)
# device-resident timings of every op (tools/quick_time.py) per variant, twice, then the GPU parity suite on the LAST variant.
set -u
lg=$1; shift
out=gpurun_out/ab_libs.log; : > $out
for rep in 1 2; do for v in "$@"; do
  echo "=== lib${v}" >> $out
  SB200_LIB=$PWD/schnorr_b200/libschnorr_b200${v}.so python tools/quick_time.py $((1<<lg)) >> $out 2>&1
done; done
last="${@: -1}"
echo "=== pytest lib${last}" >> $out
SB200_LIB=$PWD/schnorr_b200/libschnorr_b200${last}.so timeout 1500 python -m pytest tests -m gpu -x -q >> $out 2>&1
grep -E "===|verify |sign |passed|failed" $out

#!/bin/bash
# A/B timing of kernel-variant builds: tools/ab_time.sh <log2n> build/var/lib_*.so
n=$1; shift
for so in "$@"; do
  echo "== $so"
  SB200_LIB=$PWD/$so python tools/quick_time.py $n 2>&1 | grep -E "^(verify|sign|keygen)" | awk '{printf "%s %s; ", $1, $5} END {print ""}'
done

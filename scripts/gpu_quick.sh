#!/bin/bash
# GPU box: timing + stage profile only, for the product library or the variants named as arguments
mkdir -p gpurun_out
if [ $# -eq 0 ]; then
  timeout 300 python scripts/quick_bench.py 64 > gpurun_out/quick64.log 2>&1; echo "--- product rc=$?"; tail -22 gpurun_out/quick64.log
fi
for v in "$@"; do
  AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so timeout 300 python scripts/quick_bench.py 64 > gpurun_out/quick64_$v.log 2>&1
  echo "--- variant $v rc=$?"; tail -22 gpurun_out/quick64_$v.log
done

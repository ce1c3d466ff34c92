#!/bin/bash
# GPU box, short form: timing of product vs one variant library, full -m gpu suite on the variant only
mkdir -p gpurun_out
v=$1
timeout 200 python scripts/quick_bench.py 64 > gpurun_out/quick64.log 2>&1; echo "--- product B=64 rc=$?"; grep -E "^denoise|step cycles" gpurun_out/quick64.log
for b in 64 32; do
  AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so timeout 200 python scripts/quick_bench.py $b > gpurun_out/quick${b}_$v.log 2>&1; echo "--- $v B=$b rc=$?"; grep -E "^denoise|step cycles" gpurun_out/quick${b}_$v.log
done
AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -1 gpurun_out/pytest_$v.log

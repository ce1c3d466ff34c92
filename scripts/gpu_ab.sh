#!/bin/bash
# GPU box: product vs one variant library -- timing (2 clips and 1 clip per cluster) and the full -m gpu suite on both
mkdir -p gpurun_out
v=$1
for b in 64 32; do
  timeout 200 python scripts/quick_bench.py $b > gpurun_out/quick$b.log 2>&1; echo "--- product B=$b rc=$?"; grep -E "^denoise|step cycles" gpurun_out/quick$b.log
  AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so timeout 200 python scripts/quick_bench.py $b > gpurun_out/quick${b}_$v.log 2>&1; echo "--- $v B=$b rc=$?"; grep -E "^denoise|step cycles" gpurun_out/quick${b}_$v.log
done
AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -1 gpurun_out/pytest_$v.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest product rc=$?"; tail -1 gpurun_out/pytest_gpu.log

#!/bin/bash
# GPU box: parity tests with the product library, then timing + stage profile of it and of each
# variant library given as argument (names as built by scripts/build_variant.sh)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "parity\] (ddim|ddpm|backward|live ddim10 B=64|train)|passed|failed|Error|error" gpurun_out/pytest_gpu.log | tail -24
timeout 300 python scripts/quick_bench.py 64 > gpurun_out/quick64.log 2>&1; echo "--- product rc=$?"; tail -17 gpurun_out/quick64.log
for v in "$@"; do
  AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so timeout 300 python scripts/quick_bench.py 64 > gpurun_out/quick64_$v.log 2>&1
  echo "--- variant $v rc=$?"; tail -17 gpurun_out/quick64_$v.log
done

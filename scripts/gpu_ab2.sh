#!/bin/bash
# GPU box: DSMEM exchange microbenchmark, then timing + parity of each variant library given as argument
# (names as built by scripts/build_variant.sh) next to the product library.
mkdir -p gpurun_out
t0=$SECONDS
timeout 120 scripts/ubench_dsmem > gpurun_out/ubench_dsmem.log 2>&1; echo "ubench_dsmem rc=$? t=$((SECONDS-t0))s"; cat gpurun_out/ubench_dsmem.log
timeout 200 python scripts/quick_bench.py 64 > gpurun_out/quick64.log 2>&1; echo "--- product rc=$? t=$((SECONDS-t0))s"; grep -E "^denoise|step cycles" gpurun_out/quick64.log
for v in "$@"; do
  export AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so
  timeout 200 python scripts/quick_bench.py 64 > gpurun_out/quick64_$v.log 2>&1
  echo "--- variant $v rc=$? t=$((SECONDS-t0))s"; grep -E "^denoise|step cycles" gpurun_out/quick64_$v.log
  timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -1 gpurun_out/pytest_$v.log
done
unset AMUSE_B200_LIB
timeout 200 python scripts/quick_bench.py 64 > gpurun_out/quick64_again.log 2>&1; echo "--- product again rc=$? t=$((SECONDS-t0))s"; grep -E "^denoise|step cycles" gpurun_out/quick64_again.log

#!/bin/bash
# GPU box: cluster occupancy probe + step profile of the one-clip-per-cluster kernel (B <= 33)
mkdir -p gpurun_out
timeout 60 scripts/occ_probe > gpurun_out/occ_probe.log 2>&1; cat gpurun_out/occ_probe.log
for b in 32 1; do
  timeout 200 python scripts/quick_bench.py $b > gpurun_out/quick$b.log 2>&1; echo "--- B=$b rc=$?"; cat gpurun_out/quick$b.log | grep -v "^diffusion\|^decode"
done

#!/bin/bash
# GPU box: hardware probes that decide the next loop-kernel design: co-resident clusters,
# DSMEM exchange cost (4-CTA today vs the 2-CTA candidate), L2 -> shared-memory weight streaming at 1/4 and 1/2
# of the weights per SM, then the stage profile of the one-clip-per-cluster kernel (B <= 33).
mkdir -p gpurun_out
nvcc=/usr/local/cuda/bin/nvcc
for u in occ_probe ubench_dsmem ubench_l2stream; do
  [ -x scripts/$u ] || $nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/$u scripts/$u.cu
  timeout 120 scripts/$u > gpurun_out/$u.log 2>&1; echo "--- $u rc=$?"; cat gpurun_out/$u.log
done
for b in 32 1; do
  timeout 200 python scripts/quick_bench.py $b > gpurun_out/quick$b.log 2>&1; echo "--- B=$b rc=$?"; cat gpurun_out/quick$b.log | grep -v "^diffusion\|^decode"
done

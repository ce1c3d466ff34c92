#!/bin/bash
# GPU box: parity tests with the pruned last layer (default), then A/B timing + stage profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python scripts/quick_bench.py 64 > gpurun_out/quick64_prune1.log 2>&1; echo "quick1 rc=$?"
AMUSE_PRUNE_LAST=0 timeout 300 python scripts/quick_bench.py 64 > gpurun_out/quick64_prune0.log 2>&1; echo "quick0 rc=$?"
grep -E "parity|passed|failed|Error|error" gpurun_out/pytest_gpu.log | tail -30
echo "--- prune=1"; cat gpurun_out/quick64_prune1.log | tail -18
echo "--- prune=0"; cat gpurun_out/quick64_prune0.log | tail -18

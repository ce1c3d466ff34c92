#!/bin/bash
# first GPU contact: smoke, parity tests, quick timings.  Everything under timeouts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 600 python scripts/quick_bench.py 64 > gpurun_out/quick64.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/quick64.log
timeout 300 python scripts/quick_bench.py 1 > gpurun_out/quick1.log 2>&1
tail -5 gpurun_out/smoke.log; tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/quick64.log | tail -25

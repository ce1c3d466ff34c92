#!/bin/bash
# iteration loop on the GPU box: parity tests (fail fast) + quick timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 600 python scripts/quick_bench.py 64 > gpurun_out/quick64.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/quick64.log
timeout 300 python scripts/quick_bench.py 1 > gpurun_out/quick1.log 2>&1
grep -E "parity|passed|failed|Error|error" gpurun_out/pytest_gpu.log | tail -40; cat gpurun_out/quick64.log | tail -22; head -6 gpurun_out/quick1.log

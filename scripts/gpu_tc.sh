#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q -s > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?"
grep -E "tc_gemm|passed|failed|Error|error|assert" gpurun_out/pytest_tc.log | tail -30

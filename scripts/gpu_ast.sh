#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ast.py -m gpu -q -s -x > gpurun_out/pytest_ast.log 2>&1; echo "pytest rc=$?"
grep -E "parity|passed|failed|rror|assert" gpurun_out/pytest_ast.log | tail -20

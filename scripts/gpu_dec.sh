#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "decode or backward or host" > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc=$?"
grep -E "parity|passed|failed|rror" gpurun_out/pytest_dec.log | tail -20
timeout 300 python scripts/quick_bench.py 64 2>&1 | grep -E "decode|diffusion"
AMUSE_DECODE_FFMA=1 timeout 300 python scripts/quick_bench.py 64 2>&1 | grep -E "decode"

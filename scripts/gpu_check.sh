#!/bin/bash
# GPU box: the round-end checks in the driver's order -- pytest -m gpu, smoke(), bench (ours)
mkdir -p gpurun_out
t0=$SECONDS
timeout 420 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$((SECONDS-t0))s"
grep -E "passed|failed|Error|wide-rows" gpurun_out/pytest_gpu.log | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS-t0))s"; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$? t=$((SECONDS-t0))s"
cat gpurun_out/bench_ours.json

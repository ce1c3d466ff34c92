#!/bin/bash
# Round-end run on the GPU box, most important evidence first: parity with the product defaults, A/B of the
# WIDE denoise-loop layout (AMUSE_WIDE_ROWS=1) incl. its own parity pass, then bench line + ncu launch list +
# one full ncu capture of the dominant kernel with whichever layout won.  Outputs under gpurun_out/.
mkdir -p gpurun_out
t0=$SECONDS
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_idle.csv 2>&1
timeout 420 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; rc_base=$?
echo "pytest(default) rc=$rc_base t=$((SECONDS-t0))s" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -5

AMUSE_WIDE_ROWS=0 timeout 200 python scripts/quick_bench.py 64 > gpurun_out/quick64_narrow.log 2>&1; echo "--- narrow rc=$? t=$((SECONDS-t0))s"; tail -16 gpurun_out/quick64_narrow.log
AMUSE_WIDE_ROWS=1 timeout 200 python scripts/quick_bench.py 64 > gpurun_out/quick64_wide.log 2>&1; echo "--- wide rc=$? t=$((SECONDS-t0))s"; tail -16 gpurun_out/quick64_wide.log
AMUSE_WIDE_ROWS=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_train_sampler.py -m gpu -x -q -s > gpurun_out/pytest_gpu_wide.log 2>&1; rc_wide=$?
echo "pytest(wide) rc=$rc_wide t=$((SECONDS-t0))s" | tee -a gpurun_out/pytest_gpu_wide.log
grep -E "passed|failed|Error" gpurun_out/pytest_gpu_wide.log | tail -5

# winner: wide only if its parity pass is green and its ddpm1000 loop time is lower
tn=$(grep -E "^denoise B=64 ddpm1000" gpurun_out/quick64_narrow.log | sed -E 's/.*: ([0-9.]+) ms.*/\1/')
tw=$(grep -E "^denoise B=64 ddpm1000" gpurun_out/quick64_wide.log | sed -E 's/.*: ([0-9.]+) ms.*/\1/')
WIN=0
if [ "$rc_wide" = "0" ] && [ -n "$tn" ] && [ -n "$tw" ] && python -c "import sys; sys.exit(0 if float('$tw') < 0.995*float('$tn') else 1)"; then WIN=1; fi
echo "narrow=$tn ms wide=$tw ms -> AMUSE_WIDE_ROWS=$WIN" | tee gpurun_out/winner.txt
export AMUSE_WIDE_ROWS=$WIN

timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$? t=$((SECONDS-t0))s"
cat gpurun_out/bench_ours.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:denoise_loop_kernel -c 1 -o gpurun_out/prof_denoise \
    python scripts/ncu_target.py denoise > gpurun_out/ncu_denoise.log 2>&1; echo "ncu denoise rc=$? t=$((SECONDS-t0))s"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --gpus 1 --steps 1 --warmup 1 --no-baselines --no-audio > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$? t=$((SECONDS-t0))s"
echo "total t=$((SECONDS-t0))s"

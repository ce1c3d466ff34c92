#!/bin/bash
# Round-end run: product library vs the variant given as $1 (timing at 2 clips and 1 clip per cluster + the full
# -m gpu suite on both), then bench line + full ncu capture + launch list with whichever won.
mkdir -p gpurun_out
v=$1
t0=$SECONDS
timeout 300 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; rc_p=$?; echo "pytest product rc=$rc_p t=$((SECONDS-t0))s"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -1
AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so timeout 300 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_$v.log 2>&1; rc_v=$?; echo "pytest $v rc=$rc_v t=$((SECONDS-t0))s"; grep -E "passed|failed" gpurun_out/pytest_$v.log | tail -1
for b in 64 32; do
  timeout 200 python scripts/quick_bench.py $b > gpurun_out/quick$b.log 2>&1; echo "--- product B=$b rc=$?"; grep -E "^denoise|step cycles" gpurun_out/quick$b.log
  AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so timeout 200 python scripts/quick_bench.py $b > gpurun_out/quick${b}_$v.log 2>&1; echo "--- $v B=$b rc=$?"; grep -E "^denoise|step cycles" gpurun_out/quick${b}_$v.log
done
tp=$(grep -E "^denoise B=64 ddpm1000" gpurun_out/quick64.log | sed -E 's/.*: ([0-9.]+) ms.*/\1/')
tv=$(grep -E "^denoise B=64 ddpm1000" gpurun_out/quick64_$v.log | sed -E 's/.*: ([0-9.]+) ms.*/\1/')
sp=$(grep -E "^denoise B=32 ddpm1000" gpurun_out/quick32.log | sed -E 's/.*: ([0-9.]+) ms.*/\1/')
sv=$(grep -E "^denoise B=32 ddpm1000" gpurun_out/quick32_$v.log | sed -E 's/.*: ([0-9.]+) ms.*/\1/')
WIN=product
if [ "$rc_v" = "0" ] && python -c "import sys; sys.exit(0 if (float('$tv') < 0.997*float('$tp') and float('$sv') < 1.003*float('$sp')) else 1)"; then WIN=$v; export AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_$v.so; fi
echo "B=64: product=$tp $v=$tv | B=32: product=$sp $v=$sv -> winner $WIN" | tee gpurun_out/winner.txt
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$? t=$((SECONDS-t0))s"
cat gpurun_out/bench_ours.json | cut -c1-700
timeout 300 ncu --set full --clock-control none --import-source on -k regex:denoise_loop_kernel -c 1 -o gpurun_out/prof_denoise -f \
    python scripts/ncu_target.py denoise > gpurun_out/ncu_denoise.log 2>&1; echo "ncu denoise rc=$? t=$((SECONDS-t0))s"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --gpus 1 --steps 1 --warmup 1 --no-baselines --no-audio > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$? t=$((SECONDS-t0))s"

CS=/usr/local/cuda/bin/compute-sanitizer
for tool in racecheck synccheck memcheck; do
  out=gpurun_out/sanitizer_${tool}_tc_loop.txt
  timeout 600 $CS --tool $tool --print-limit 20 python scripts/sanitize_target.py denoise > $out 2>&1
  echo "--- $tool tc_loop rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard' $out | tail -2 | tr '\n' ' ')"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python scripts/quick_bench.py 64 2>&1 | head -18

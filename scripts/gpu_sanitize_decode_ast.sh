CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  for tgt in decode ast; do
    out=gpurun_out/sanitizer_${tool}_${tgt}.txt
    timeout 500 $CS --tool $tool --print-limit 20 python scripts/sanitize_target.py $tgt > $out 2>&1
    echo "--- $tool $tgt rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard' $out | tail -2 | tr '\n' ' ')"
  done
done

#!/bin/bash
# GPU box: compute-sanitizer memcheck / synccheck / racecheck over every kernel family (SURVEY.md section 5 "race
# detection"): the tcgen05 sampler loop (mbarrier / st.async / TMEM protocols), the FFMA fallback loop in both layouts
# (barrier-free DSMEM exchange, WIDE buffer aliasing), one decode, one AST depth-1 pass, the filterbank.
# Summaries land in gpurun_out/sanitizer_<tool>_<target>.txt (copied to profiles/r02_sanitizer_*.txt).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # tool target [env...]
  local tool=$1 tgt=$2 tag=$3; shift 3
  local out=gpurun_out/sanitizer_${tool}_${tag}.txt
  env "$@" timeout 600 $CS --tool $tool --print-limit 20 python scripts/sanitize_target.py $tgt > $out 2>&1
  echo "--- $tool $tag rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard' $out | tail -2 | tr '\n' ' ')"
}
for tool in memcheck synccheck racecheck; do
  run $tool denoise tc_loop AMUSE_DENOISE_FFMA=0
  run $tool denoise ffma_loop AMUSE_DENOISE_FFMA=1
  run $tool denoise ffma_loop_wide AMUSE_DENOISE_FFMA=1 AMUSE_WIDE_ROWS=1
  run $tool decode decode X=1
  run $tool ast ast X=1
  run $tool fbank fbank X=1
done

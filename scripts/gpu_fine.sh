#!/bin/bash
# GPU box: inside-stage cycle stamps (library built with -DAMUSE_FINE_PROF) for 2 clips and 1 clip per cluster,
# then the bench line of the product library
mkdir -p gpurun_out
for b in 64 32; do
  AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_fine.so timeout 200 python scripts/quick_bench.py $b > gpurun_out/fine$b.log 2>&1
  echo "--- fine B=$b rc=$?"; grep -E "^denoise|^fine|step cycles" gpurun_out/fine$b.log
done
timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_ours.json')); print(d['value'], d['ms_per_step'], d['roofline'])"

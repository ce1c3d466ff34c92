#!/bin/bash
# Round-end evidence run on the GPU box: bench lines (ours + reference arm), ncu launch list of the
# bench command, one full ncu capture of the dominant kernel, clocks.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_idle.csv 2>&1
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
# launch list (every kernel with its device time) of a short bench run
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --gpus 1 --steps 1 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
# full capture of the dominant kernel (the persistent loop) on a SHORT schedule so the ~40 replays stay cheap
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:denoise_loop_kernel -c 1 -o gpurun_out/prof_denoise \
    python scripts/ncu_target.py denoise > gpurun_out/ncu_denoise.log 2>&1; echo "ncu denoise rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 10 -c 3 -o gpurun_out/prof_decode_gemm \
    python scripts/ncu_target.py decode > gpurun_out/ncu_decode.log 2>&1; echo "ncu decode rc=$?"
cat gpurun_out/bench_ours.json; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ours.err

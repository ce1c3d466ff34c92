#!/bin/bash
# GPU box, round 2 evidence run: full -m gpu suite, smoke, bench (ours + reference arm), ncu launch list of the bench
# command, full ncu capture of the dominant kernel (the tcgen05 sampler loop) on a short schedule.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --gpus 1 --steps 1 --warmup 1 --no-baselines > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:denoise_tc_kernel -c 1 -f -o gpurun_out/prof_dtc \
    python scripts/ncu_target.py denoise > gpurun_out/ncu_dtc.log 2>&1; echo "ncu denoise rc=$?"
AMUSE_DECODE_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|self_attention" -s 12 -c 8 -f \
    -o gpurun_out/prof_decode python scripts/ncu_target.py decode > gpurun_out/ncu_decode.log 2>&1; echo "ncu decode rc=$?"
cat gpurun_out/bench_ours.json | cut -c1-1500; echo; cat gpurun_out/bench_ref.json

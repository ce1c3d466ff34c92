#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 8 -c 4 -o gpurun_out/prof_tc_gemm_ast python scripts/ast_bench.py 16 > gpurun_out/ncu_tc.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file gpurun_out/launches_ast.csv python scripts/ast_bench.py 16 > gpurun_out/ncu_ast_list.log 2>&1; echo rc=$?
tail -2 gpurun_out/ncu_tc.log

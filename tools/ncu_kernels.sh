#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:gather_gemm_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r2_gg32_subm_sorted python tools/prof_layer.py 2 1 sorted 2>&1 | grep -E "sorted|rror"
timeout 300 $NCU -k regex:gather_gemm_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r2_gg64_subm_sorted python tools/prof_layer.py 3 1 sorted 2>&1 | grep -E "sorted|rror"
timeout 300 $NCU -k regex:gather_gemm_tc_kernel --launch-skip 3 -c 1 -o gpurun_out/r2_gg128_subm_sorted python tools/prof_layer.py 4 1 sorted 2>&1 | grep -E "sorted|rror"
timeout 300 $NCU -k regex:gather_wgrad_rows -c 1 -o gpurun_out/r2_wgrad32 python tools/prof_layer.py 2 1 wgrad 2>&1 | grep -E "wgrad|rror"
timeout 300 $NCU -k regex:gather_wgrad_rows -c 1 -o gpurun_out/r2_wgrad64 python tools/prof_layer.py 3 1 wgrad 2>&1 | grep -E "wgrad|rror"
ls -la gpurun_out/*.ncu-rep
echo "== prof_step"; timeout 300 python tools/prof_step.py > gpurun_out/prof_step.txt 2>&1; head -50 gpurun_out/prof_step.txt

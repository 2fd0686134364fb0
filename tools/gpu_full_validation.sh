#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q -rf 2>&1 | grep -vE "^\s*$|Warning|warn|run_backward|Consider|Docs|meshgrid|_VF" | tail -12 | cut -c1-500 | tee gpurun_out/pytest_gpu_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:gather_gemm_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r2b_gg32_subm_sorted python tools/prof_layer.py 2 1 sorted 2>&1 | grep -E "sorted|rror"
timeout 300 $NCU -k regex:gather_wgrad_rows -c 1 -o gpurun_out/r2b_wgrad32 python tools/prof_layer.py 2 1 wgrad 2>&1 | grep -E "wgrad|rror"
timeout 300 $NCU -k regex:gather_wgrad_rows -c 1 -o gpurun_out/r2b_wgrad128 python tools/prof_layer.py 4 1 wgrad 2>&1 | grep -E "wgrad|rror"
echo "== layers"; for s in 1 2 3 4 5 6; do timeout 120 python tools/prof_layer.py $s 5; done 2>&1 | tee gpurun_out/layers_final.txt

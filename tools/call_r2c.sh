#!/bin/bash
# round-2 GPU call C: dense TMA kernel tests first (bounded), then the suite, layer timings, ncu captures, bench
set -u
mkdir -p gpurun_out
echo "== conv2d tests"; timeout 300 python -m pytest tests/test_gpu_conv2d.py -m gpu -q -x -rf 2>&1 | grep -vE "^\s*$" | tail -25 | tee gpurun_out/pytest_conv2d.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -rf --deselect tests/test_gpu_conv2d.py 2>&1 | grep -vE "^\s*$" | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== layers"; for s in 1 2 3 4 5 6; do timeout 120 python tools/prof_layer.py $s 5; done 2>&1 | tee gpurun_out/layers.txt
echo "== ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -c 1 -f -o gpurun_out/r2_gg16_sorted python tools/prof_layer.py 1 1 sorted 2>&1 | grep -E "sorted|rror" 
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -c 1 -f -o gpurun_out/r2_gg32_sorted python tools/prof_layer.py 2 1 sorted 2>&1 | grep -E "sorted|rror"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -c 1 -f -o gpurun_out/r2_conv2d_tma256 python tools/prof_layer.py 5 1 tma 2>&1 | grep -E "tma|rror"
ls -la gpurun_out/*.ncu-rep | tail -4
echo "== bench"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
tail -3 gpurun_out/bench.err; head -24 gpurun_out/groups.txt

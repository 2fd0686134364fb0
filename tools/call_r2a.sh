#!/bin/bash
# round-2 GPU call A: full parity suite (all failures, not -x), gather4 micro-benchmark, bench, glue attribution
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
echo "== gather4 micro"; timeout 120 tools/micro/gather4_bench 2>&1 | tee gpurun_out/gather4_bench.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -s -rf 2>&1 | grep -vE "^\s*$" | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== bench"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-600
tail -3 gpurun_out/bench.err
echo "== glue"; timeout 400 python tools/prof_glue.py > gpurun_out/prof_glue.txt 2>&1; head -50 gpurun_out/prof_glue.txt

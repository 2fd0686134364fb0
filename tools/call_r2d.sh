#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x -rf 2>&1 | grep -vE "^\s*$" | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== dense BN knob"; for s in 5 6; do CPD_DENSE_BN=128 timeout 120 python tools/prof_layer.py $s 5 tma; done 2>&1 | tee gpurun_out/layers_bn128.txt
echo "== bench"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
tail -3 gpurun_out/bench.err; head -12 gpurun_out/groups.txt; tail -5 gpurun_out/groups.txt
echo "== bench prefetch"; timeout 900 python bench.py --steps 10 --warmup 3 --prefetch --no-cpu-baseline 2> gpurun_out/bench_prefetch.err | tee gpurun_out/bench_prefetch.json | cut -c1-200
echo "== stress"; CPD_BENCH_GROUPS=gpurun_out/groups_stress.txt timeout 900 python bench.py --workload stress --steps 10 --warmup 3 2> gpurun_out/bench_stress.err | tee gpurun_out/bench_stress.json | cut -c1-300
tail -3 gpurun_out/bench_stress.err; cat gpurun_out/groups_stress.txt
echo "== glue"; timeout 400 python tools/prof_glue.py > gpurun_out/prof_glue.txt 2>&1; head -4 gpurun_out/prof_glue.txt

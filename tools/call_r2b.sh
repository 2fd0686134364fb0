#!/bin/bash
# round-2 GPU call B: parity suite + bench (+ optional extra command in $EXTRA)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -s -rf ${PYTEST_ARGS:-} 2>&1 | grep -vE "^\s*$" | tail -${PYTEST_TAIL:-40} | tee gpurun_out/pytest_gpu.log
echo "== bench"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
tail -3 gpurun_out/bench.err
head -30 gpurun_out/groups.txt
if [ -n "${EXTRA:-}" ]; then echo "== extra"; bash -c "$EXTRA"; fi

#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_conv2d.py -m gpu -q -x -rf 2>&1 | grep -vE "^\s*$|Warning|warn" | tail -12 | cut -c1-300 | tee gpurun_out/pytest_parity.log
echo "== layers"; for s in 1 2 3 4; do timeout 120 python tools/prof_layer.py $s 5; done 2>&1 | tee gpurun_out/layers.txt
echo "== bench prefetch"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prefetch 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json")); print(round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s; launches", d["gpu_launches"])
PY
head -30 gpurun_out/groups.txt

#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== roipool + graph tests"; timeout 900 python -m pytest tests/test_gpu_roipool.py tests/test_gpu_detector.py -m gpu -q -rf 2>&1 | grep -vE "^\s*$|Warning|warn|run_backward|Consider|Docs" | tail -30 | tee gpurun_out/pytest_roipool.log
echo "== bench (graph)"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
echo "== bench (graph + prefetch)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prefetch 2> gpurun_out/bench_prefetch.err | tee gpurun_out/bench_prefetch.json | cut -c1-200
python - <<'PY'
import json
for f in ("bench","bench_prefetch"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s; launches", d["gpu_launches"])
    except Exception as e: print(f, "failed", e)
PY

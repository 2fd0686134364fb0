#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== graph test"; timeout 600 python -m pytest tests/test_gpu_detector.py -m gpu -q -x -rf -k "graph or prepared or train_step" 2>&1 | grep -vE "^\s*$" | tail -25 | tee gpurun_out/pytest_graph.log
echo "== prof_step"; timeout 300 python tools/prof_step.py 2>&1 | head -4 | tee gpurun_out/prof_step_head.txt
echo "== bench (graph)"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-250
tail -3 gpurun_out/bench.err
echo "== bench (no graph)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph 2> gpurun_out/bench_nograph.err | tee gpurun_out/bench_nograph.json | cut -c1-250
echo "== bench (graph + prefetch)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prefetch 2> gpurun_out/bench_prefetch.err | tee gpurun_out/bench_prefetch.json | cut -c1-250
tail -3 gpurun_out/bench_prefetch.err
python - <<'PY'
import json
for f in ("bench","bench_nograph","bench_prefetch"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s; launches", d["gpu_launches"])
    except Exception as e: print(f, "failed", e)
PY

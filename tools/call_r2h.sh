#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | grep -vE "^\s*$|Warning|warn|run_backward|Consider|Docs" | tail -25 | cut -c1-400 | tee gpurun_out/pytest_gpu_all.log
echo "== graph diag"; timeout 600 python tools/diag_graph.py 2>&1 | grep -vE "Warning|warn|run_backward|Consider" | tee gpurun_out/diag_graph.txt
for cfg in "--no-graph" "--no-graph --prefetch" "" "--prefetch"; do
  tag=$(echo "bench$cfg" | tr -d ' ' | tr '-' '_')
  echo "== bench $cfg"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg 2> gpurun_out/$tag.err > gpurun_out/$tag.json
  python - "$tag" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s; launches", d["gpu_launches"])
except Exception as e: print(f, "failed", e)
PY
done

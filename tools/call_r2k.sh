#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== layers"; for s in 2 3; do timeout 120 python tools/prof_layer.py $s 5 wgrad; done 2>&1 | tee gpurun_out/layers_wgrad.txt
for cfg in "" "--no-graph" "" "--no-graph"; do
  echo "== bench $cfg"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg 2> gpurun_out/bench.err > gpurun_out/bench.json
  python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json")); print(round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"])
PY
done

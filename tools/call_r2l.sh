#!/bin/bash
set -u
mkdir -p gpurun_out
for h in 0 1; do
echo "== layers CPD_L2_HINTS=$h"; for s in 1 2 3 4; do CPD_L2_HINTS=$h timeout 120 python tools/prof_layer.py $s 5; done 2>&1 | grep -v "^stage" | tee gpurun_out/layers_hints$h.txt
done
for h in 0 1; do
  echo "== bench CPD_L2_HINTS=$h"; CPD_L2_HINTS=$h timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_hints$h.json
  python - $h <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_hints{sys.argv[1]}.json")); print(round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"])
PY
done

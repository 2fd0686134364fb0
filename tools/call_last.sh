#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench"; timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_last.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_last.json")); print(round(d["ms_per_step"],2), "ms/step; value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"], d["allocator_rank0"], d["clocks"])
PY
echo "== pytest -m gpu"; timeout 170 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | cut -c1-300

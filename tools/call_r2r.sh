#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -rf -x 2>&1 | grep -vE "^\s*$|Warning|warn|run_backward|Consider|Docs" | tail -8 | cut -c1-400 | tee gpurun_out/pytest_gpu_all.log
for i in 1 2; do
echo "== bench $i"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json")); print(round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"], "launches", d["gpu_launches"])
PY
done

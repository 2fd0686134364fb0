#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_detector.py tests/test_gpu_parity2.py tests/test_gpu_roipool.py -m gpu -q -rf 2>&1 | grep -vE "^\s*$|Warning|warn|run_backward|Consider|Docs" | tail -8 | cut -c1-400
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for i in 1 2 3 4; do
echo "== bench $i"; CPD_BENCH_DIAG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench$i.err > gpurun_out/bench.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json")); print(round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"], d["allocator_rank0"])
PY
done
grep diag gpurun_out/bench1.err | cut -c1-160

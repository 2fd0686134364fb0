#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== wgrad parity"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_conv2d.py -m gpu -q -x -rf -k "wgrad or gather or conv or bench_scale or detector" 2>&1 | grep -vE "^\s*$|Warning|warn|Consider" | tail -5 | cut -c1-300
for z in 0 1; do
echo "== layers wgrad CPD_WGRAD_SKIP_ZERO=$z"; for s in 1 2 3 4; do CPD_WGRAD_SKIP_ZERO=$z timeout 120 python tools/prof_layer.py $s 5 wgrad; done 2>&1 | grep -v "^stage"
done | tee gpurun_out/layers_wgrad3.txt
echo "== bench 30 steps diag"; CPD_BENCH_DIAG=1 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_diag.err > gpurun_out/bench.json
grep diag gpurun_out/bench_diag.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json")); print(round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s;", "peak GB", d["peak_hbm_gb_rank0"])
PY

#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== layers"; for s in 1 2 3 4; do timeout 120 python tools/prof_layer.py $s 5 wgrad; done 2>&1 | tee gpurun_out/layers_wgrad2.txt
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | grep -vE "^\s*$|Warning|warn|run_backward|Consider|Docs" | tail -12 | cut -c1-400 | tee gpurun_out/pytest_gpu_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

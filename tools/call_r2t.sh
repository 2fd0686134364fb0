#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_properties.py tests/test_gpu_roipool.py -m gpu -q -rf 2>&1 | grep -vE "^\s*$|Warning|warn|run_backward|Consider|Docs" | tail -12 | cut -c1-500

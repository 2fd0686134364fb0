#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== graph test"; timeout 600 python -m pytest tests/test_gpu_detector.py -m gpu -q -x -rf -k "graph" 2>&1 | grep -vE "^\s*$" | tail -12 | tee gpurun_out/pytest_graph.log
echo "== diag"; timeout 600 python tools/diag_e2e.py --graph 2>&1 | grep -v "Warning\|warn\|run_backward" | tee gpurun_out/diag_e2e.txt
echo "== diag expandable"; timeout 600 python tools/diag_e2e.py --graph --expandable 2>&1 | grep -v "Warning\|warn\|run_backward" | tee gpurun_out/diag_e2e_exp.txt

#!/bin/bash
set -u
mkdir -p gpurun_out
for d in 0 2 4 8 10 14; do
echo "== CPD_TC_DEBUG=$d"; for s in 1 2 3 4; do CPD_TC_DEBUG=$d timeout 120 python tools/prof_layer.py $s 5 sorted; done 2>&1 | grep -v "^stage"
done | tee gpurun_out/layers_debug2.txt

#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -q -x -rf 2>&1 | grep -vE "^\s*$|Warning|warn" | tail -6 | cut -c1-300
echo "== bench"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -2 gpurun_out/bench.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json")); print(round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"])
PY
tail -5 gpurun_out/groups.txt
echo "== prof_step"; timeout 300 python tools/prof_step.py > gpurun_out/prof_step.txt 2>&1; grep -E "strided|subm_table|bitmap|hash_insert|bn_bwd_reduce|bn_bwd_apply|bn_apply|tap_block|tile_masks|aten::sort|aten::gather|aten::index_select|aten::copy_" gpurun_out/prof_step.txt | cut -c1-60,150-230 | head -30

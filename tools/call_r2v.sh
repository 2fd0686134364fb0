#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests (detector, anchor head)"; timeout 900 python -m pytest tests/test_gpu_detector.py tests/test_gpu_anchor_head.py -m gpu -q -rf 2>&1 | grep -vE "^\s*$|Warning|warn|run_backward|Consider|Docs|meshgrid|_VF" | tail -6 | cut -c1-400
echo "== bench default (with cpu baseline)"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench.err > gpurun_out/bench_final.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final.json")); print(round(d["ms_per_step"],2), "ms/step; value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"], d["allocator_rank0"], d.get("cpu_baseline"))
PY
echo "== bench backbone_fwd"; timeout 600 python bench.py --workload backbone_fwd --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_bb.err > gpurun_out/bench_backbone_fwd.json
echo "== bench stress"; CPD_BENCH_GROUPS=gpurun_out/groups_stress.txt timeout 600 python bench.py --workload stress --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_stress.err > gpurun_out/bench_stress.json
python - <<'PY'
import json
for f in ("bench_backbone_fwd","bench_stress"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"],2), "ms/step; value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
    except Exception as e: print(f, "failed", e)
PY
cat gpurun_out/groups_stress.txt

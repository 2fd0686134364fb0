#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["ms_per_step"],2), "ms/step; value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"], d["allocator_rank0"])
except Exception as e: print("failed", e)
PY
}
for i in 1 2 3; do
echo "== bench default $i"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench.json; show gpurun_out/bench.json
done
echo "== bench prefetch-thread"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --prefetch-thread 2> gpurun_out/bench_t.err > gpurun_out/bench_thread.json; show gpurun_out/bench_thread.json
echo "== stress"; timeout 600 python bench.py --workload stress --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_stress.err > gpurun_out/bench_stress.json; show gpurun_out/bench_stress.json
echo "== backbone"; timeout 600 python bench.py --workload backbone_fwd --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_bb.err > gpurun_out/bench_backbone_fwd.json; show gpurun_out/bench_backbone_fwd.json

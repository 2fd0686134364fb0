#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d["ms_per_step"],2), "ms/step; value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"], d["allocator_rank0"], d["clocks"])
except Exception as e: print("failed", e)
PY
}
for i in 1 2 3; do
echo "== bench default $i"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench.json; show gpurun_out/bench.json
done
for i in 1 2; do
echo "== bench prefetch-thread $i"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --prefetch-thread 2> gpurun_out/bench_t.err > gpurun_out/bench_thread.json; show gpurun_out/bench_thread.json; tail -2 gpurun_out/bench_t.err | cut -c1-200
done
echo "== stress"; timeout 600 python bench.py --workload stress --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_stress.err > gpurun_out/bench_stress.json; show gpurun_out/bench_stress.json
echo "== prepared test"; timeout 600 python -m pytest tests/test_gpu_detector.py -m gpu -q -k prepared 2>&1 | tail -2

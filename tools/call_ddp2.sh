#!/bin/bash
# usage: gpurun --gpus 2 -- 'bash tools/call_ddp2.sh'   -- DDP option experiments at N=2
set -u
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/ddp_$tag.err > gpurun_out/ddp_$tag.json
  python - $tag <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ddp_{t}.json").read().strip().splitlines()[-1]); print(t, round(d["ms_per_step"],2), "ms/step;", d["ms_steps_rank0"])
except Exception as e: print(t, "failed", e)
PY
}
run default
CPD_DDP_KWARGS='{"broadcast_buffers": false}' run nobcast
CPD_DDP_KWARGS='{"gradient_as_bucket_view": true}' run bucketview
CPD_DDP_KWARGS='{"broadcast_buffers": false, "gradient_as_bucket_view": true, "static_graph": true}' run all3
run nograph --no-graph

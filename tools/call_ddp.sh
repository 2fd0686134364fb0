#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/call_ddp.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
nproc; nvidia-smi -L | head -8
echo "== bench N=$N"
CPD_BENCH_DIAG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_EXTRA:-} 2> gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1]); print("N=",n, round(d["ms_per_step"],2), "ms/step; value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"])
except Exception as e: print("failed", e)
PY
grep "\[diag\]" gpurun_out/bench_n$N.err | head -12 | cut -c1-170

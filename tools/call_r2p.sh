#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
echo "== bench $i"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json")); print(round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "frames/s;", d["ms_steps_rank0"], "peak GB", d["peak_hbm_gb_rank0"], "launches", d["gpu_launches"])
PY
done
echo "== ncu launch list (default command: prefetch + graphs)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5600 -c 3400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
wc -l gpurun_out/r2_launches.csv

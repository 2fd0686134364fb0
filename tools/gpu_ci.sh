#!/bin/bash
# Run on the B200 box via gpurun: parity tests, smoke, a short bench (+ optional kernel/launch profiles).
#   gpurun --timeout 1500 -- 'bash tools/gpu_ci.sh'                 tests + smoke + bench
#   gpurun --timeout 1500 -- 'PROFILE=1 bash tools/gpu_ci.sh'       + torch-profiler kernel table (tools/prof_step.py)
#   tools/call_prof.sh adds the ncu launch list (slow: ncu serialises ~1500 launches per step);
#   tools/call_final.sh captures `ncu --set full` reports of the top kernels with tools/prof_layer.py;
#   read reports here with tools/ncu_summary.py, tools/ncu_stalls.py, tools/ncu_lines.py.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ "${PROFILE:-0}" = "1" ]; then
  echo "== torch profiler"; timeout 300 python tools/prof_step.py > gpurun_out/prof_step.txt 2>&1; head -3 gpurun_out/prof_step.txt
fi

set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
echo "== prof"; timeout 300 python tools/prof_step.py > gpurun_out/prof_step.txt 2>&1; head -5 gpurun_out/prof_step.txt

set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|shared conv" | tail -8
echo "== bench"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
echo "== prof"; timeout 300 python tools/prof_step.py > gpurun_out/prof_step.txt 2>&1; head -3 gpurun_out/prof_step.txt; grep -E "^void cpd|^cpd::|Memcpy|Memset|at::native" gpurun_out/prof_step.txt | head -40 | cut -c1-60,150-215

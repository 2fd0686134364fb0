set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench"; CPD_BENCH_GROUPS=gpurun_out/groups.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
echo "== ncu"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -s 3 -c 1 -f -o gpurun_out/f_gg128 python tools/prof_layer.py 4 1 2>&1 | grep -E "gather_gemm:|error" 
timeout 250 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -s 2 -c 1 -f -o gpurun_out/f_gg64 python tools/prof_layer.py 3 1 2>&1 | grep -E "gather_gemm:|error"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:gather_wgrad_rows -s 1 -c 1 -f -o gpurun_out/f_wg64 python tools/prof_layer.py 3 1 2>&1 | grep -E "tap-major|error"
ls -la gpurun_out/*.ncu-rep | tail -4

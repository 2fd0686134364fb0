set -u
mkdir -p gpurun_out
echo "== prof"; timeout 300 python tools/prof_step.py > gpurun_out/prof_step.txt 2>&1; head -3 gpurun_out/prof_step.txt
echo "== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
wc -l gpurun_out/launches.csv

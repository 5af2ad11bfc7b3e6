#!/bin/bash
# One GPU session: all parity tests, smoke, profile, bench + ncu launch list. Logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu --timeout 400 -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1
echo "all_gpu_tests exit $?" > gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "profile exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary.txt
NSLAM_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -25 gpurun_out/all_gpu_tests.log
tail -16 gpurun_out/profile_step.log
cat gpurun_out/bench.json
tail -5 gpurun_out/bench.err
wc -l gpurun_out/launches.csv

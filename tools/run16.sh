#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/t16.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 300 python tools/host_profile.py > gpurun_out/host_profile.log 2>&1; echo "hostprof exit $?" >> gpurun_out/summary.txt
NSLAM_UPDATE_GRAPHS=0 timeout 300 python tools/host_profile.py > gpurun_out/host_profile_eager.log 2>&1; echo "hostprof eager exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
NSLAM_UPDATE_GRAPHS=0 timeout 300 python bench.py --steps 64 --warmup 8 > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "bench eager exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/t16.log; head -16 gpurun_out/host_profile.log | cut -c1-160;  head -16 gpurun_out/host_profile_eager.log | cut -c1-160; cut -c1-330 gpurun_out/bench.json; echo; cut -c1-330 gpurun_out/bench_eager.json; tail -3 gpurun_out/bench.err

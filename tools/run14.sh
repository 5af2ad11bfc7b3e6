#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ngp.py tests/test_gpu_frontend.py tests/test_gpu_glue.py -m gpu -q > gpurun_out/t14.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 400 python tools/host_profile.py > gpurun_out/host_profile.log 2>&1; echo "hostprof exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -8 gpurun_out/t14.log; head -30 gpurun_out/host_profile.log | cut -c1-160; cut -c1-500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err

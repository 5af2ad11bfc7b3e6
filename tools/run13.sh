#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t13.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 400 python tools/host_profile.py > gpurun_out/host_profile.log 2>&1; echo "hostprof exit $?" >> gpurun_out/summary.txt
NERF_ITERS=0 timeout 400 python tools/host_profile.py > gpurun_out/host_profile_nonerf.log 2>&1; echo "hostprof0 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -30 gpurun_out/t13.log; head -70 gpurun_out/host_profile.log; head -3 gpurun_out/host_profile_nonerf.log

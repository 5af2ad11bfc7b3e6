#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t15.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 400 python tools/host_profile.py > gpurun_out/host_profile.log 2>&1; echo "hostprof exit $?" >> gpurun_out/summary.txt
timeout 400 python tools/kernel_table.py > gpurun_out/kernel_table.log 2>&1; echo "ktable exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -8 gpurun_out/t15.log; head -24 gpurun_out/host_profile.log | cut -c1-160; grep "^==" gpurun_out/kernel_table.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err

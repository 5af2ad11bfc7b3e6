#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ngp.py -m gpu -x -q > gpurun_out/t9.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 400 python tools/kernel_table.py > gpurun_out/kernel_table.log 2>&1; echo "ktable exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -25 gpurun_out/t9.log; cat gpurun_out/kernel_table.log | cut -c1-220; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

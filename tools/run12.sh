#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t12.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 400 python tools/kernel_table.py > gpurun_out/kernel_table.log 2>&1; echo "ktable exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -30 gpurun_out/t12.log; grep -v "^{" gpurun_out/kernel_table.log | head; cut -c1-600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err

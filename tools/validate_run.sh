#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t23.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table.log 2>&1; echo "ktable exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -n "FAILED\|passed\|failed" gpurun_out/t23.log | head; grep "^==" gpurun_out/kernel_table.log; grep "im2col\|ba_solve" gpurun_out/kernel_table.log | cut -c1-140; cut -c1-330 gpurun_out/bench.json; tail -n 1 gpurun_out/smoke.log

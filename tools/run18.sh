#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t18.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table.log 2>&1; echo "ktable exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -6 gpurun_out/t18.log; grep "^==" gpurun_out/kernel_table.log; grep "conv_igemm" gpurun_out/kernel_table.log | head -12 | cut -c1-150; cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/smoke.log

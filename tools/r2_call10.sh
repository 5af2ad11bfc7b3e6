#!/bin/bash
# Round-2 GPU call 10: convolutions with four TMA-issuing warps — parity first, then kernel tables and the bench
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -x tests/test_gpu_conv.py tests/test_gpu_golden.py tests/test_gpu_glue.py > gpurun_out/t10_conv.log 2>&1; echo "conv tests exit $?" > gpurun_out/summary.txt
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table10.log 2>&1
timeout 500 python bench.py > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t10_conv.log
grep -h "== \|conv_" gpurun_out/kernel_table10.log | cut -c1-150 | head -44
cut -c1-300 gpurun_out/bench10.json; tail -2 gpurun_out/bench10.err

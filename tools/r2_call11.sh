#!/bin/bash
# Round-2 GPU call 11: CTA-pair convolution kernel with four TMA-issuing warps — parity, kernel table, bench (all with NSLAM_CONV_CTA2=1)
mkdir -p gpurun_out
export NSLAM_CONV_CTA2=1
timeout 600 python -m pytest -q -m gpu -x tests/test_gpu_conv.py tests/test_gpu_golden.py tests/test_gpu_glue.py > gpurun_out/t11_conv.log 2>&1; echo "conv tests (pairs) exit $?" > gpurun_out/summary.txt
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table11_pairs.log 2>&1
timeout 500 python bench.py > gpurun_out/bench11_pairs.json 2> gpurun_out/bench11.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t11_conv.log
grep -h "== \|conv_" gpurun_out/kernel_table11_pairs.log | cut -c1-150 | head -30
cut -c1-300 gpurun_out/bench11_pairs.json; tail -2 gpurun_out/bench11.err

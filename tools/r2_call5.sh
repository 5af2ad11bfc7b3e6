#!/bin/bash
# Round-2 GPU call 5: second-generation 3x3 convolution (conv_halo.cu) — parity first, then timing; lookup A/B
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -x tests/test_gpu_conv.py tests/test_gpu_golden.py tests/test_gpu_shapes.py -k "conv or update or encoder or golden" > gpurun_out/t5_conv.log 2>&1; echo "conv tests exit $?" > gpurun_out/summary.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/t5.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table5.log 2>&1
NSLAM_CONV_GEN1=1 NSLAM_LOOKUP_SCALAR=1 timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table5_gen1.log 2>&1
timeout 500 python bench.py > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 15 gpurun_out/t5_conv.log; tail -n 12 gpurun_out/t5.log
grep -h "== update\|conv_\|corr_lookup" gpurun_out/kernel_table5.log | cut -c1-150 | head -16
echo GEN1; grep -h "== update\|conv_\|corr_lookup" gpurun_out/kernel_table5_gen1.log | cut -c1-150 | head -14
cut -c1-400 gpurun_out/bench5.json; tail -3 gpurun_out/bench5.err

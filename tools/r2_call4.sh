#!/bin/bash
# Round-2 GPU call 4: warp-uniform MMA/TMA issue (conv, pairs, NeRF), 128-bit lookup, volume epilogue, slot build
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/t4.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
NSLAM_CONV_CTA2=1 timeout 600 python -m pytest -q -m gpu tests/test_gpu_conv.py > gpurun_out/t4_pairs.log 2>&1; echo "pairs tests exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table4.log 2>&1
NSLAM_CONV_CTA2=1 timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table4_pairs.log 2>&1
timeout 500 python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench exit $?" >> gpurun_out/summary.txt
NSLAM_E=16 timeout 200 python tools/microbench.py 2> /dev/null > gpurun_out/microbench4.jsonl
cat gpurun_out/summary.txt; tail -n 12 gpurun_out/t4.log; tail -n 4 gpurun_out/t4_pairs.log
grep -h "== update\|conv_igemm\|corr_lookup\|corr_volume\|== frame\|== context\|== nerf\|backward_tc\|forward_tc" gpurun_out/kernel_table4.log | cut -c1-160 | head -40
echo PAIRS; grep -h "== update\|conv_igemm" gpurun_out/kernel_table4_pairs.log | cut -c1-160 | head -12
cut -c1-900 gpurun_out/bench4.json; head -5 gpurun_out/microbench4.jsonl | cut -c1-250

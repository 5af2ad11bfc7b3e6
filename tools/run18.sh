#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t18.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table.log 2>&1; echo "ktable exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
NSLAM_ENCODER=cudnn timeout 300 python bench.py --steps 64 --warmup 8 > gpurun_out/bench_cudnn_enc.json 2> gpurun_out/bench_cudnn_enc.err; echo "bench cudnn-enc exit $?" >> gpurun_out/summary.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -n "FAILED\|passed\|failed\|Error" gpurun_out/t18.log | head -20; grep "^==" gpurun_out/kernel_table.log; grep "conv_igemm\|ba_solve" gpurun_out/kernel_table.log | head -24 | cut -c1-150; cut -c1-400 gpurun_out/bench.json; echo; cut -c1-300 gpurun_out/bench_cudnn_enc.json; echo; tail -n 2 gpurun_out/smoke.log

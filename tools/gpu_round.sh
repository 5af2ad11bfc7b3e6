#!/bin/bash
# One GPU session: parity tests (each file in its own bounded process), micro-benchmarks, smoke.
# Everything is logged under gpurun_out/ so a cut-off call can still be read.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
export PYTHONUNBUFFERED=1
for t in tests/test_gpu_parity.py tests/test_gpu_vs_reference.py; do
  n=$(basename $t .py)
  timeout 600 python -m pytest $t -q -m gpu --timeout 180 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit $?" >> gpurun_out/summary.txt
done
true
true
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -60 gpurun_out/test_gpu_parity.log
tail -40 gpurun_out/test_gpu_vs_reference.log
cat gpurun_out/microbench.jsonl
tail -3 gpurun_out/microbench.err
timeout 600 python -m pytest tests/test_gpu_frontend.py -q -m gpu --timeout 500 -p no:cacheprovider > gpurun_out/test_gpu_frontend.log 2>&1
echo "frontend exit $?" >> gpurun_out/summary.txt
tail -40 gpurun_out/test_gpu_frontend.log
N=150 timeout 600 python tools/run_slam.py > gpurun_out/run_slam.log 2>&1
echo "run_slam exit $?" >> gpurun_out/summary.txt
tail -45 gpurun_out/run_slam.log

#!/bin/bash
# Round-2 GPU call 14 (4 GPUs): bench at N=4 as the driver launches it (3 data-parallel trainers: gradient all-reduce in their group)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n4.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 4 --steps 192 --warmup 8 > gpurun_out/bench14_n4.json 2> gpurun_out/bench14_n4.err; echo "bench N=4 exit $?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt; cut -c1-1300 gpurun_out/bench14_n4.json; tail -n 6 gpurun_out/bench14_n4.err

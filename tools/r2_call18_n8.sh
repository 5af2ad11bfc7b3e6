#!/bin/bash
# Round-2 GPU call 18 (8 GPUs): bench at N=8 as the driver launches it (7 data-parallel trainers: gradient all-reduce in their group)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n8.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 8 --steps 192 --warmup 8 > gpurun_out/bench18_n8.json 2> gpurun_out/bench18_n8.err; echo "bench N=8 exit $?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt; cut -c1-1300 gpurun_out/bench18_n8.json; tail -n 6 gpurun_out/bench18_n8.err

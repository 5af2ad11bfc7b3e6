#!/bin/bash
# Round-2 GPU call 7 (2 GPUs): 2-rank NCCL equality / hand-off test, then the bench at N=2 as the driver launches it
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n2.txt 2>&1
timeout 500 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/t7_dist.log 2>&1; echo "dist test exit $?" > gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 192 --warmup 8 > gpurun_out/bench7_n2.json 2> gpurun_out/bench7_n2.err; echo "bench N=2 exit $?" >> gpurun_out/summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench7_n2_ref.json 2> gpurun_out/bench7_n2_ref.err; echo "ref arm N=2 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 8 gpurun_out/t7_dist.log; cut -c1-1500 gpurun_out/bench7_n2.json; tail -n 8 gpurun_out/bench7_n2.err; cut -c1-400 gpurun_out/bench7_n2_ref.json; tail -n 3 gpurun_out/bench7_n2_ref.err

#!/bin/bash
# multi-GPU bench: $1 = number of GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 96 --warmup 8 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?" > gpurun_out/summary_n$N.txt
cat gpurun_out/summary_n$N.txt; cut -c1-1200 gpurun_out/bench_n$N.json; tail -n 8 gpurun_out/bench_n$N.err

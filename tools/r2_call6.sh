#!/bin/bash
# Round-2 GPU call 6: tensor-pipe / L2 fill probes, TSDF test, cfg4 / cfg5 workloads, bench on the default kernels
mkdir -p gpurun_out
timeout 200 python tools/probes/run_umma_rate_probe.py > gpurun_out/umma_rate.log 2>&1; echo "probe exit $?" > gpurun_out/summary.txt
timeout 300 python -m pytest -q -m gpu tests/test_gpu_tsdf.py tests/test_gpu_parity.py -k "tsdf or lookup or ba_" > gpurun_out/t6.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 500 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 700 python bench.py --workload cfg5 --steps 96 > gpurun_out/bench6_cfg5.json 2> gpurun_out/bench6_cfg5.err; echo "cfg5 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --workload cfg4 --steps 480 > gpurun_out/bench6_cfg4.json 2> gpurun_out/bench6_cfg4.err; echo "cfg4 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/umma_rate.log; tail -n 6 gpurun_out/t6.log
for f in bench6 bench6_cfg5 bench6_cfg4; do echo $f; cut -c1-500 gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done

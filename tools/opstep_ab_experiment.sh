#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_frontend.py tests/test_gpu_golden.py -m gpu -q > gpurun_out/t21.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
: > gpurun_out/ab.log
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 192 --warmup 8 2> gpurun_out/bench_$label.err | tee gpurun_out/bench_$label.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$label', d['value'], 'fps e2e', d['e2e']['value'], 'kf', d['config']['keyframes_in_timed_region'], 'upd', d['config']['update_calls_in_timed_region'])" >> gpurun_out/ab.log 2>&1
}
run opstep_eager NSLAM_UPDATE_GRAPHS=0
run opstep_graph NSLAM_UPDATE_GRAPHS=1
run python_graph NSLAM_OP_STEP=0 NSLAM_UPDATE_GRAPHS=1
run opstep_eager_cudnnenc NSLAM_UPDATE_GRAPHS=0 NSLAM_ENCODER=cudnn
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table.log 2>&1; echo "ktable exit $?" >> gpurun_out/summary.txt
NSLAM_TIMERS=1 NSLAM_CPROFILE=0 NSLAM_UPDATE_GRAPHS=0 timeout 300 python tools/host_profile.py > gpurun_out/host_timers.log 2>&1
cat gpurun_out/summary.txt; grep -n "FAILED\|passed\|failed" gpurun_out/t21.log | head; cat gpurun_out/ab.log; grep "^==" gpurun_out/kernel_table.log; grep "conv_igemm" gpurun_out/kernel_table.log | head -12 | cut -c1-150; head -2 gpurun_out/host_timers.log | cut -c1-1300

#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sm_split.log
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 64 --warmup 8 2> gpurun_out/bench_$label.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$label', d['value'], 'fps e2e', d['e2e']['value'])" >> gpurun_out/sm_split.log 2>&1
}
run tc_148_148 NSLAM_ENCODER=tcgen05
run tc_108_40 NSLAM_ENCODER=tcgen05 NSLAM_SLAM_SMS=108 NSLAM_NERF_SMS=40
run tc_116_32 NSLAM_ENCODER=tcgen05 NSLAM_SLAM_SMS=116 NSLAM_NERF_SMS=32
run tc_148_40 NSLAM_ENCODER=tcgen05 NSLAM_NERF_SMS=40
run cudnn_148_40 NSLAM_ENCODER=cudnn NSLAM_NERF_SMS=40
run cudnn_116_32 NSLAM_ENCODER=cudnn NSLAM_SLAM_SMS=116 NSLAM_NERF_SMS=32
NSLAM_TIMERS=1 NSLAM_CPROFILE=0 NSLAM_ENCODER=cudnn timeout 300 python tools/host_profile.py > gpurun_out/host_timers.log 2>&1
cat gpurun_out/sm_split.log; head -3 gpurun_out/host_timers.log | cut -c1-1500

#!/bin/bash
# Round-2 GPU call 2: full ungated GPU suite, product bench, reference-CUDA arm, host profile, targeted ncu captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/t2.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 500 python bench.py > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 500 python bench.py --impl reference-cuda --steps 96 --warmup 8 > gpurun_out/refcuda2.json 2> gpurun_out/refcuda2.err; echo "refcuda exit $?" >> gpurun_out/summary.txt
NSLAM_TIMERS=1 NSLAM_CPROFILE=1 timeout 300 python tools/host_profile.py > gpurun_out/host_profile2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"conv_igemm_kernel|corr_lookup_nhwc|corr_volume|backward_tc_kernel|forward_tc_kernel|grid_scatter" \
   -o /tmp/r02_call2 -f python tools/ncu_targets.py > gpurun_out/ncu2.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
ncu -i /tmp/r02_call2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_call2.csv 2>/dev/null
for k in "conv_igemm_kernel<128, 2" "conv_igemm_kernel<16, 0" "conv_igemm_kernel<256, 1" "corr_lookup_nhwc" "corr_volume" "backward_tc_kernel"; do
  f=$(echo "$k" | tr -c 'a-zA-Z0-9' '_')
  python tools/ncu_hot_lines.py /tmp/r02_call2.ncu-rep "$k" 45 0 > gpurun_out/r02_hotlines_$f.txt 2>&1
done
ls -la /tmp/r02_call2.ncu-rep >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 25 gpurun_out/t2.log; cut -c1-1500 gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err; cat gpurun_out/refcuda2.json; tail -5 gpurun_out/refcuda2.err; head -3 gpurun_out/host_profile2.log

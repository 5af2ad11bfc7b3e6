#!/bin/bash
# Round-2 GPU call 3: suite with the new rows (TSDF, pose refinement, render parity, demo wiring, async prefetch ...),
# bench after the host-side changes, UMMA descriptor probe, ncu --set full of the update operator's convolutions
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/t3.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 500 python bench.py > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 200 python tools/probes/run_umma_probe.py > gpurun_out/umma_probe.log 2>&1; echo "probe exit $?" >> gpurun_out/summary.txt
NSLAM_TIMERS=1 NSLAM_CPROFILE=1 timeout 300 python tools/host_profile.py > gpurun_out/host_profile3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"conv_igemm_kernel" -c 16 -o /tmp/r02_call3 -f python tools/ncu_targets.py > gpurun_out/ncu3.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
ncu -i /tmp/r02_call3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_call3.csv 2>/dev/null
for k in "conv_igemm_kernel<128, 2" "conv_igemm_kernel<16, 0" "conv_igemm_kernel<256, 1" "conv_igemm_kernel<128, 0, true"; do
  f=$(echo "$k" | tr -c 'a-zA-Z0-9' '_')
  python tools/ncu_hot_lines.py /tmp/r02_call3.ncu-rep "$k" 60 0 > gpurun_out/r02_hotlines_$f.txt 2>&1
done
ls -la /tmp/r02_call3.ncu-rep >> gpurun_out/summary.txt
sz=$(stat -c %s /tmp/r02_call3.ncu-rep); if [ "$sz" -lt 45000000 ]; then cp /tmp/r02_call3.ncu-rep gpurun_out/; fi
cat gpurun_out/summary.txt; tail -n 25 gpurun_out/t3.log; cut -c1-700 gpurun_out/bench3.json; tail -3 gpurun_out/bench3.err; head -60 gpurun_out/umma_probe.log; head -3 gpurun_out/host_profile3.log

#!/bin/bash
# Round-2 GPU call 12: ncu --set full of every kernel of one update(), one frame front and one NeRF step (default kernels),
# source-level stall samples of the update operator's convolutions
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -o /tmp/r02_call12 -f python tools/ncu_targets.py > gpurun_out/ncu12.log 2>&1; echo "ncu exit $?" > gpurun_out/summary.txt
ncu -i /tmp/r02_call12.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_call12.csv 2>/dev/null
ncu -i /tmp/r02_call12.ncu-rep --page details --csv > gpurun_out/r02_ncu_details_call12.csv 2>/dev/null
for i in $(seq 0 14); do
  python tools/ncu_hot_lines.py /tmp/r02_call12.ncu-rep "conv_igemm_kernel" 28 $i > gpurun_out/r02_hotlines_conv_launch$i.txt 2>&1
done
python tools/ncu_hot_lines.py /tmp/r02_call12.ncu-rep "corr_lookup_nhwc_kernel" 25 0 > gpurun_out/r02_hotlines_corr_lookup.txt 2>&1
python tools/ncu_hot_lines.py /tmp/r02_call12.ncu-rep "backward_tc_kernel" 25 0 > gpurun_out/r02_hotlines_backward_tc.txt 2>&1
python tools/ncu_hot_lines.py /tmp/r02_call12.ncu-rep "cam_grad_kernel" 25 0 > gpurun_out/r02_hotlines_cam_grad.txt 2>&1
ls -la /tmp/r02_call12.ncu-rep >> gpurun_out/summary.txt
sz=$(stat -c %s /tmp/r02_call12.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/r02_call12.ncu-rep gpurun_out/; fi
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/ncu12.log; wc -l gpurun_out/r02_ncu_raw_call12.csv

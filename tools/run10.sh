#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_ngp_bwd.py > gpurun_out/debug_bwd.log 2>&1; echo "debug exit $?" > gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"backward_tc_kernel|forward_tc_kernel|conv_igemm_kernel|corr_volume_tc_kernel|sample_rays|loss_kernel|cvx_upsample|corr_lookup_nhwc|ba_solve|ba_linearize|ba_schur_kernel" \
   -o gpurun_out/r01_full -f python tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
ls -la gpurun_out/*.ncu-rep >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/debug_bwd.log | tail -40; tail -15 gpurun_out/ncu_full.log

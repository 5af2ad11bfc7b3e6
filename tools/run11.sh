#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_ngp_bwd.py > gpurun_out/debug_bwd.log 2>&1; echo "debug exit $?" > gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_ngp.py -m gpu -q > gpurun_out/t11.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"backward_tc_kernel|forward_tc_kernel|conv_igemm_kernel|sample_rays|loss_kernel|cvx_upsample|corr_lookup_nhwc|ba_solve|ba_linearize|ba_schur_kernel" \
   -o /tmp/r01_full -f python tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
ls -la /tmp/r01_full.ncu-rep >> gpurun_out/summary.txt
ncu -i /tmp/r01_full.ncu-rep --page raw --csv > gpurun_out/r01_ncu_raw.csv 2>/dev/null
ncu -i /tmp/r01_full.ncu-rep --page details --csv > gpurun_out/r01_ncu_details.csv 2>/dev/null
sz=$(stat -c %s /tmp/r01_full.ncu-rep); if [ "$sz" -lt 55000000 ]; then cp /tmp/r01_full.ncu-rep gpurun_out/; fi
NSLAM_E=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"corr_volume_tc_kernel" -s 2 -c 1 -o gpurun_out/r01_corrvol -f python tools/microbench.py > gpurun_out/ncu_cv.log 2>&1; echo "ncu cv exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
du -sh gpurun_out >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/debug_bwd.log | tail -32; tail -12 gpurun_out/t11.log; tail -3 gpurun_out/ncu_full.log; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err

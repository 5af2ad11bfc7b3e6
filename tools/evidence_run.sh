#!/bin/bash
# final evidence run: full GPU tests, ncu --set full of one update() + one frame front + one NeRF step, launch list, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t22.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"backward_tc_kernel|forward_tc_kernel|grid_scatter|conv_igemm_kernel|sample_rays|loss_kernel|cvx_upsample|corr_lookup_nhwc|corr_volume_tc|ba_solve|ba_linearize|ba_schur_kernel|motion_im2col|inorm_apply|im2col7" \
   -o /tmp/r01_final -f python tools/ncu_targets.py > gpurun_out/ncu_final.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
ncu -i /tmp/r01_final.ncu-rep --page raw --csv > gpurun_out/r01_ncu_raw_final.csv 2>/dev/null
ncu -i /tmp/r01_final.ncu-rep --page details --csv > gpurun_out/r01_ncu_details_final.csv 2>/dev/null
ls -la /tmp/r01_final.ncu-rep >> gpurun_out/summary.txt
sz=$(stat -c %s /tmp/r01_final.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/r01_final.ncu-rep gpurun_out/; fi
# launch list (durations only) of a short bench run
NSLAM_CUDA_PROFILER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r01_ncu_launches_final.csv python bench.py --steps 12 --warmup 4 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo "launch list exit $?" >> gpurun_out/summary.txt
NSLAM_E=16 timeout 200 python tools/microbench.py > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
du -sh gpurun_out >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -n "FAILED\|passed\|failed" gpurun_out/t22.log | head; tail -n 3 gpurun_out/ncu_final.log; head -4 gpurun_out/microbench.jsonl | cut -c1-250; wc -l gpurun_out/r01_ncu_launches_final.csv; cat gpurun_out/bench.json

#!/bin/bash
# Round-2 GPU call 8: where a frame's time goes at one GPU (CUPTI timeline, with and without the NeRF trainer), TMA tensor-load
# fill rate of the activation boxes, halo convolution after the producer reordering
mkdir -p gpurun_out
timeout 200 python tools/probes/run_umma_rate_probe.py --boxes-only > gpurun_out/box_fill.log 2>&1; echo "probe exit $?" > gpurun_out/summary.txt
timeout 400 python tools/timeline.py 48 2 > gpurun_out/timeline_nerf2.log 2>&1; echo "timeline exit $?" >> gpurun_out/summary.txt
timeout 400 python tools/timeline.py 48 0 > gpurun_out/timeline_nerf0.log 2>&1; echo "timeline0 exit $?" >> gpurun_out/summary.txt
NSLAM_CONV_HALO=1 timeout 300 python -m pytest -q -m gpu -x tests/test_gpu_conv.py > gpurun_out/t8_conv_halo.log 2>&1; echo "halo conv tests exit $?" >> gpurun_out/summary.txt
NSLAM_CONV_HALO=1 timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table8_halo.log 2>&1
cat gpurun_out/summary.txt; cat gpurun_out/box_fill.log; grep -v Warn gpurun_out/timeline_nerf2.log | cut -c1-200; echo ----; grep -v Warn gpurun_out/timeline_nerf0.log | cut -c1-200 | head -30
tail -n 3 gpurun_out/t8_conv_halo.log; grep -h "== update\|conv_halo\|conv_igemm" gpurun_out/kernel_table8_halo.log | cut -c1-150 | head -14

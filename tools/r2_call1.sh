#!/bin/bash
# Round-2 GPU call 1: run everything that was gated/opt-in at the end of round 1.
mkdir -p gpurun_out
export NSLAM_PENDING_TESTS=1
timeout 600 python -m pytest -q -m gpu \
  "tests/test_gpu_ngp.py::test_sample_rays_matches_the_march_oracle" \
  "tests/test_gpu_ngp.py::test_process_slam_ingest_matches_reference_golden" \
  "tests/test_gpu_parity.py::test_ba_covariances_reference_exact" \
  "tests/test_gpu_parity.py::test_droid_backends_ba_all_in_one_loop" > gpurun_out/pending_tests.log 2>&1
echo "pending tests exit $?" > gpurun_out/summary.txt
timeout 900 python -m pytest -q -m gpu tests/test_gpu_droid.py > gpurun_out/pending_droid.log 2>&1
echo "pending droid.py tests exit $?" >> gpurun_out/summary.txt
NSLAM_CORRVOL_ROWS=1 timeout 300 python -m pytest -q -m gpu tests/test_gpu_parity.py -k "corr_volume" > gpurun_out/corr_rows_tests.log 2>&1
echo "corr rows tests exit $?" >> gpurun_out/summary.txt
NSLAM_CONV_CTA2=1 timeout 600 python -m pytest -q -m gpu tests/test_gpu_conv.py > gpurun_out/conv_pairs_tests.log 2>&1
echo "conv pairs tests exit $?" >> gpurun_out/summary.txt
NSLAM_CONV_CTA2=1 timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table_pairs.log 2>&1
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table_default.log 2>&1
timeout 300 python examples/slam_demo.py --dataset_dir=synthetic --dataset_name=nerf --buffer=60 --slam --fusion=nerf --synthetic_frames 120 > gpurun_out/demo_synthetic.log 2>&1
echo "slam_demo synthetic exit $?" >> gpurun_out/summary.txt
NSLAM_TIMERS=1 NSLAM_CPROFILE=0 timeout 300 python tools/host_profile.py > gpurun_out/host_timers.log 2>&1
NSLAM_E=16 timeout 200 python tools/microbench.py 2> /dev/null | head -3 > gpurun_out/microbench_tiled.jsonl
NSLAM_CORRVOL_ROWS=1 NSLAM_E=16 timeout 200 python tools/microbench.py 2> /dev/null | head -3 > gpurun_out/microbench_rows.jsonl
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 30 gpurun_out/pending_tests.log; tail -n 30 gpurun_out/pending_droid.log; tail -n 8 gpurun_out/corr_rows_tests.log; tail -n 12 gpurun_out/conv_pairs_tests.log
grep -h "conv_igemm" gpurun_out/kernel_table_pairs.log | head -8; grep -h "conv_igemm" gpurun_out/kernel_table_default.log | head -8
head -n 2 gpurun_out/host_timers.log; tail -n 3 gpurun_out/demo_synthetic.log; head -1 gpurun_out/microbench_tiled.jsonl; head -1 gpurun_out/microbench_rows.jsonl; cut -c1-400 gpurun_out/bench_default.json

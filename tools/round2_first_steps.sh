#!/bin/bash
# First GPU call of the next round: run what was written at the end of round 1 without GPU access.
#   1. the gated tests (ray sampler vs march oracle, droid_backends.ba, the DROID-style plugin surface droid.py)
#   2. the experimental row-pair correlation-volume kernel: bit-exactness vs the tiled kernel, micro-benchmark of
#      both, the whole GPU suite and the bench with it enabled
mkdir -p gpurun_out
NSLAM_PENDING_TESTS=1 timeout 600 python -m pytest -q -m gpu \
  "tests/test_gpu_ngp.py::test_sample_rays_matches_the_march_oracle" \
  "tests/test_gpu_ngp.py::test_process_slam_ingest_matches_reference_golden" \
  "tests/test_gpu_parity.py::test_ba_covariances_reference_exact" \
  "tests/test_gpu_parity.py::test_droid_backends_ba_all_in_one_loop" > gpurun_out/pending_tests.log 2>&1
echo "pending tests exit $?" > gpurun_out/summary.txt
NSLAM_PENDING_TESTS=1 timeout 900 python -m pytest -q -m gpu tests/test_gpu_droid.py > gpurun_out/pending_droid.log 2>&1
echo "pending droid.py tests exit $?" >> gpurun_out/summary.txt
NSLAM_CORRVOL_ROWS=1 timeout 300 python -m pytest -q -m gpu tests/test_gpu_parity.py -k "corr_volume" > gpurun_out/corr_rows_tests.log 2>&1
echo "corr rows tests exit $?" >> gpurun_out/summary.txt
# 3. the CTA-pair convolution kernel (csrc/conv_igemm2.cu, cta_group::2): parity, then the kernel table with it enabled
NSLAM_CONV_CTA2=1 timeout 600 python -m pytest -q -m gpu tests/test_gpu_conv.py > gpurun_out/conv_pairs_tests.log 2>&1
echo "conv pairs tests exit $?" >> gpurun_out/summary.txt
NSLAM_CONV_CTA2=1 timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -m gpu tests/test_gpu_conv.py -k "pairs or gru_fused" > gpurun_out/conv_pairs_memcheck.log 2>&1
echo "conv pairs memcheck exit $?" >> gpurun_out/summary.txt
NSLAM_CONV_CTA2=1 timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table_pairs.log 2>&1
timeout 300 python tools/kernel_table.py > gpurun_out/kernel_table_default.log 2>&1
# 3b. the reference's command line on this implementation (procedural stream, then the same stream from files)
timeout 300 python examples/slam_demo.py --dataset_dir=synthetic --dataset_name=nerf --buffer=60 --slam --fusion=nerf --synthetic_frames 120 > gpurun_out/demo_synthetic.log 2>&1
echo "slam_demo synthetic exit $?" >> gpurun_out/summary.txt
python -c "from nerf_slam_b200 import datasets, synthetic; datasets.write_transforms_dataset(synthetic.SyntheticRoom(640, 480, 60), '/tmp/nslam_ds')" \
  && timeout 300 python examples/slam_demo.py --dataset_dir=/tmp/nslam_ds --dataset_name=nerf --buffer=60 --slam --fusion=nerf --eval > gpurun_out/demo_files.log 2>&1
echo "slam_demo files exit $?" >> gpurun_out/summary.txt
# 3c. host-side timers after the native graph / proximity routines (compare with profiles/r01_host_timers_run21.log)
NSLAM_TIMERS=1 NSLAM_CPROFILE=0 timeout 300 python tools/host_profile.py > gpurun_out/host_timers.log 2>&1
# 4. hardware question for the next convolution redesign (one halo box for all nine taps): see tools/probes/
timeout 200 python tools/probes/run_umma_probe.py > gpurun_out/umma_probe.log 2>&1
NSLAM_E=16 timeout 200 python tools/microbench.py 2> /dev/null | head -3 > gpurun_out/microbench_tiled.jsonl
NSLAM_CORRVOL_ROWS=1 NSLAM_E=16 timeout 200 python tools/microbench.py 2> /dev/null | head -3 > gpurun_out/microbench_rows.jsonl
NSLAM_CORRVOL_ROWS=1 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/suite_with_rows.log 2>&1
echo "suite with rows kernel exit $?" >> gpurun_out/summary.txt
NSLAM_CORRVOL_ROWS=1 timeout 400 python bench.py > gpurun_out/bench_rows.json 2> gpurun_out/bench_rows.err
echo "bench with rows kernel exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 15 gpurun_out/pending_tests.log; tail -n 25 gpurun_out/pending_droid.log; tail -n 5 gpurun_out/corr_rows_tests.log; tail -n 8 gpurun_out/conv_pairs_tests.log; grep -h "conv_igemm" gpurun_out/kernel_table_pairs.log | head -8; grep -h "conv_igemm" gpurun_out/kernel_table_default.log | head -8
head -n 2 gpurun_out/host_timers.log; tail -n 3 gpurun_out/demo_synthetic.log; tail -n 3 gpurun_out/demo_files.log; cat gpurun_out/umma_probe.log | head -40; head -1 gpurun_out/microbench_tiled.jsonl; head -1 gpurun_out/microbench_rows.jsonl; tail -n 3 gpurun_out/suite_with_rows.log; cut -c1-300 gpurun_out/bench_rows.json

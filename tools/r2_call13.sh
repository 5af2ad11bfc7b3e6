#!/bin/bash
# Round-2 GPU call 13 (evidence run on the final default kernels): full GPU suite, smoke, launch list of the bench command,
# micro-benchmarks, bench (ours / reference-cuda / reference)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/t13.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke13.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
NSLAM_CUDA_PROFILER=1 timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_ncu_launches_bench_call13.csv python bench.py --steps 12 --warmup 4 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo "launch list exit $?" >> gpurun_out/summary.txt
NSLAM_E=16 timeout 200 python tools/microbench.py > gpurun_out/microbench13.jsonl 2> gpurun_out/microbench13.err; echo "microbench exit $?" >> gpurun_out/summary.txt
timeout 500 python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 500 python bench.py --impl reference-cuda > gpurun_out/bench13_refcuda.json 2> gpurun_out/bench13_refcuda.err; echo "reference-cuda exit $?" >> gpurun_out/summary.txt
timeout 500 python bench.py --impl reference --steps 12 --warmup 1 > gpurun_out/bench13_ref.json 2> gpurun_out/bench13_ref.err; echo "reference exit $?" >> gpurun_out/summary.txt
du -sh gpurun_out >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 12 gpurun_out/t13.log; tail -2 gpurun_out/smoke13.log; wc -l gpurun_out/r02_ncu_launches_bench_call13.csv
cut -c1-400 gpurun_out/bench13.json; cut -c1-400 gpurun_out/bench13_refcuda.json; cut -c1-400 gpurun_out/bench13_ref.json; grep corr_volume gpurun_out/microbench13.jsonl | cut -c1-250

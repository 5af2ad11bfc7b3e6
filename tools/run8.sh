mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_ngp.py tests/test_gpu_parity.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/t8.log
echo "tests exit $?" > gpurun_out/summary.txt
timeout 400 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "profile exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py 2>gpurun_out/microbench.err | head -3 > gpurun_out/microbench.jsonl
timeout 500 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/t8.log; tail -16 gpurun_out/profile_step.log; cat gpurun_out/microbench.jsonl; cat gpurun_out/bench.json; tail -12 gpurun_out/bench.err

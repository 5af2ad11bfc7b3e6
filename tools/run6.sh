mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -40 | tee gpurun_out/test_gpu_conv.log

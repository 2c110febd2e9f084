python -m pytest tests/test_gpu_fused.py tests/test_gpu_extra.py -x -q 2>&1 | tail -15
python tools/fused_bench.py 32768 2>&1 | tail -12

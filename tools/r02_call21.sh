python -m pytest tests/test_gpu_sharded_abi.py tests/test_gpu_fused.py -x -q 2>&1 | tail -12

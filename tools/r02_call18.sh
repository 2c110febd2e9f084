python -m pytest tests/test_gpu_reproducible.py tests/test_gpu_mirror.py -x -q 2>&1 | tail -8
python tools/matvec_k_bench.py 262144
python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -x -q 2>&1 | tail -4

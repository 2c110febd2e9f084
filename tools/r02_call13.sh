python tools/matvec_k_bench.py 262144
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4

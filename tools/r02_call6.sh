python -m pytest tests/test_gpu_gram_wide.py -x -q 2>&1 | tail -3
python tools/gram_wide_bench.py 16384 17 32 64 100 256 2>&1 | grep "ident=on"

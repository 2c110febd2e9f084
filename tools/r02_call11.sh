python -m pytest tests/test_gpu_reproducible.py tests/test_gpu_parity.py -x -q 2>&1 | tail -12
python tools/prof_driver.py matvec 32768 1048576 2>&1 | tail -3

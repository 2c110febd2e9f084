set -x
python -m pytest tests/test_gpu_gram_wide.py -x -q > gpurun_out/r02_gramwide_test.log 2>&1; tail -30 gpurun_out/r02_gramwide_test.log
python tools/gram_wide_bench.py 16384 > gpurun_out/r02_gramwide_bench.log 2>&1; cat gpurun_out/r02_gramwide_bench.log

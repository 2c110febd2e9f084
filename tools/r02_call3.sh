set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest3.log 2>&1; tail -5 gpurun_out/r02_gputest3.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/r02_bench3.json
tail -20 gpurun_out/r02_bench3.err

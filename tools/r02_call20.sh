set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest20.log 2>&1; tail -4 gpurun_out/r02_gputest20.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench20.json 2> gpurun_out/r02_bench20.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench20_ref.json 2> gpurun_out/r02_bench20_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-callers --no-cpu-baseline > gpurun_out/r02_ncu_launch.log 2>&1
tail -c 1500 gpurun_out/r02_bench20.json

python -m pytest tests/test_gpu_mirror.py -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-matvec --no-callers > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench14.json'))
print(json.dumps(d['e2e'])[:900]); print(d['value'], d['extra']['per_kernel_ms'])
PY
nproc; lscpu | grep -i "model name\|socket\|numa\|^cpu(s)"

python -m pytest tests/test_gpu_reproducible.py tests/test_gpu_sharded_abi.py -x -q 2>&1 | tail -6
python tools/check_sharded_nccl.py > /dev/null 2>&1 || true
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_g2b.json 2> gpurun_out/r02_bench_g2b.err; echo "rc=$?"
tail -3 gpurun_out/r02_bench_g2b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_g2b.json'))
print(d.get('parity_failed'), d['value'], json.dumps(d['e2e'])[:500])
ex=d['extra']
for k in ('matvec','matvec_k8','problinsolve'):
    print(k, json.dumps(ex.get(k))[:900])
PY

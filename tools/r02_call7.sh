set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_g2.json 2> gpurun_out/r02_bench_g2.err; echo "rc=$?"
tail -5 gpurun_out/r02_bench_g2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_g2.json'))
print(d.get('parity_failed'), d['value'], d['e2e'])
ex=d['extra']
for k in ('gram_row_sharded','matvec','matvec_k8','problinsolve'):
    print(k, json.dumps(ex.get(k))[:1500])
PY

set -x
nproc; lscpu | grep -i "socket\|numa\|^cpu(s)\|model name"; free -g | head -2
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 tools/host_bw_probe.py --gib 4 --numa >> gpurun_out/r02_hostbw_multi.json 2>> gpurun_out/r02_hostbw_multi.err
done
cat gpurun_out/r02_hostbw_multi.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_g8.json 2> gpurun_out/r02_bench_g8.err; echo "rc=$?"
tail -3 gpurun_out/r02_bench_g8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_g8.json'))
print(d.get('parity_failed'), d['value'], json.dumps(d['e2e'])[:600])
ex=d['extra']
for k in ('gram_row_sharded','matvec','matvec_k8','problinsolve'):
    print(k, json.dumps(ex.get(k))[:1300])
PY

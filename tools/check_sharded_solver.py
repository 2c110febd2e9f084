"""BASELINE.json configs[3] on several GPUs: the reference's problinsolve on k.linop(X) + 1e-3 I with the
drop-in installed in sharded mode (one process per GPU, products spread over the ranks, one NCCL
collective per product).  Checks the iterate against the single-GPU run and prints timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29521 tools/check_sharded_solver.py [n_check] [n_full]
"""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import probnum_b200.patch as gpu  # noqa: E402

gpu.numpy2_compat()
warnings.simplefilter("ignore")
import probnum as pn  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["PNCOV_DEVICE"] = str(local)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_check = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n_full = int(sys.argv[2]) if len(sys.argv) > 2 else 65536


def solve(n, sharded, maxiter=50):
    rng = np.random.default_rng(42)
    X = rng.normal(size=(n, 4))
    y = rng.normal(size=n)
    k = pn.randprocs.covfuncs.ExpQuad((4,), lengthscales=0.25)
    gpu.enable(sharded=sharded)
    try:
        t0 = time.perf_counter()
        L = k.linop(X)
        x, _, _, info = pn.linalg.problinsolve(L + 1e-3 * pn.linops.Identity(n), y, maxiter=maxiter)
        dt = time.perf_counter() - t0
        kind = type(L).__name__
    finally:
        gpu.disable()
    return x.mean, dt, int(info["iter"]), kind


x1, t1, it1, kind1 = solve(n_check, False)
xs, ts, its, kinds = solve(n_check, True)
assert kinds == "GpuShardedCovarianceLinearOperator" and kind1 == "GpuCovarianceLinearOperator"
assert it1 == its
rel = float(np.abs(xs - x1).max() / np.abs(x1).max())
assert rel <= 1e-8, rel
xf, tf, itf, _ = solve(n_full, True)
t = torch.from_numpy(np.ascontiguousarray(xf)).cuda()
ref = t.clone()
dist.broadcast(ref, 0)
assert torch.equal(t, ref), "replicas diverged"
if rank == 0:
    print(json.dumps({"world": world, "n_check": n_check, "iter": its, "x_rel_diff_sharded_vs_single": rel,
                      "single_gpu_s": t1, "sharded_s": ts, "n_full": n_full, "sharded_full_s": tf, "iter_full": itf}))
dist.destroy_process_group()

"""Run under torchrun on >= 2 GPUs: checks ShardedCovarianceOperator (host and device operands,
NCCL all-gather) against the oracle on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tools/check_sharded_nccl.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from oracle import covfuncs_oracle as orc  # noqa: E402
from probnum_b200.sharded import ShardedCovarianceOperator  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(3)
for (n0, n1, d, kernel, spec) in [
    (4099, 3001, 4, pb.ExpQuad((4,), lengthscales=1.1), orc.KernelSpec("expquad", (4,), 1.1)),
    (4096, 4096, 3, pb.Matern((3,), nu=2.5), orc.KernelSpec("matern", (3,), 1.0, nu=2.5)),
]:
    x0, x1 = rng.normal(size=(n0, d)), rng.normal(size=(n1, d))
    v = rng.normal(size=(n1, 2))
    op = ShardedCovarianceOperator(kernel, x0, x1, device=local)
    y = op @ v                                     # host operands, uneven shards allowed
    want = orc.matvec(spec, x0, x1, v)
    bound = orc.abs_matvec(spec, x0, x1, v)
    assert np.all(np.abs(y - want) <= 1e-12 * bound), f"rank {rank}: host path mismatch"
    Kb = op.gram_block()                           # this rank's row block of the Gram matrix
    np.testing.assert_allclose(Kb, orc.gram(spec, x0[op.lo:op.hi], x1), rtol=1e-12, atol=1e-12)
    if n0 % world == 0:
        vd = torch.from_numpy(v).to(dev)
        yd = op.matmul_device(vd).cpu().numpy()     # device operands, in-place all-gather
        assert np.all(np.abs(yd - want) <= 1e-12 * bound), f"rank {rank}: device path mismatch"
    op.close()
# symmetric operators, one column: round-robin row blocks + all-reduce (pncov_matmul_sym_part_device)
for (n, d, kernel, spec) in [
    (5003, 8, pb.ExpQuad((8,)), orc.KernelSpec("expquad", (8,), 1.0)),
    (4096, 3, pb.RatQuad((3,), alpha=2.0), orc.KernelSpec("ratquad", (3,), 1.0, alpha=2.0)),
    (4096, 3, pb.Matern((3,), nu=1.5), orc.KernelSpec("matern", (3,), 1.0, nu=1.5)),  # direct-difference form
]:
    x = rng.normal(size=(n, d))
    v = rng.normal(size=(n, 1))
    op = ShardedCovarianceOperator(kernel, x, None, device=local)
    want = orc.matvec(spec, x, None, v)
    bound = orc.abs_matvec(spec, x, None, v)
    y = op @ v
    assert np.all(np.abs(y - want) <= 1e-12 * bound), f"rank {rank}: symmetric host path mismatch"
    used_sym = op.local_sym is not None
    assert used_sym, spec.kind
    if n % world == 0:
        yd = op.matmul_device(torch.from_numpy(v).to(dev)).cpu().numpy()
        assert np.all(np.abs(yd - want) <= 1e-12 * bound), f"rank {rank}: symmetric device path mismatch"
    op.close()
dist.barrier()
if rank == 0:
    print(f"sharded NCCL check ok on {world} GPUs")
dist.destroy_process_group()

"""Row-sharded operator over torch.distributed with world_size = 2 on CPU (gloo).

The per-rank compute is a stand-in built from the oracle (there is no GPU here); what is
under test is the N > 1 host path: row partition, replicated operand, all-gather and
reassembly of uneven shards, identical results on every rank."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import probnum_b200 as pb
from oracle import covfuncs_oracle as orc
from probnum_b200.sharded import ShardedCovarianceOperator, shard_bounds

SPEC = orc.KernelSpec("matern", (3,), 0.8, nu=1.5)


class OracleLocal:
    """Stand-in for _lib.Operator: same matmul() contract, NumPy arithmetic (test only)."""

    device = 0

    def __init__(self, params, rows, cols):
        assert params.d == 3 and rows.flags.c_contiguous and (cols is None or cols.flags.c_contiguous)
        self.rows, self.cols = rows, cols

    def matmul(self, v):
        return orc.matvec(SPEC, self.rows, self.cols if self.cols is not None else self.rows, v)

    gram_dtype = np.float64

    def gram(self, out):
        out[...] = orc.gram(SPEC, self.rows, self.cols if self.cols is not None else self.rows)
        return out

    def matmul_sym_part(self, v, part, nparts, block=16):
        """Contribution of row blocks part, part + nparts, ... to K @ v, each off-diagonal tile
        used for y_I and y_J (the decomposition of pncov_matmul_sym_part_device)."""
        x = self.rows
        n = x.shape[0]
        y = np.zeros((n, v.shape[1]))
        nb = (n + block - 1) // block
        for bi in range(part, nb, nparts):
            i0, i1 = bi * block, min((bi + 1) * block, n)
            y[i0:i1] += orc.gram(SPEC, x[i0:i1], x[i0:i1]) @ v[i0:i1]
            if i1 < n:
                K = orc.gram(SPEC, x[i0:i1], x[i1:])
                y[i0:i1] += K @ v[i1:]
                y[i1:] += K.T @ v[i0:i1]
        return y


def _worker(rank, world, port, n0, n1, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        x0 = rng.normal(size=(n0, 3))
        x1 = rng.normal(size=(n1, 3))
        v = rng.normal(size=(n1, 2))
        k = pb.Matern((3,), nu=1.5, lengthscales=0.8)
        op = ShardedCovarianceOperator(k, x0, x1, local_factory=OracleLocal)
        assert op.shape == (n0, n1)
        assert (op.lo, op.hi) == shard_bounds(n0, world, rank)
        y = op @ v
        y1 = op @ v[:, 0]
        want = orc.matvec(SPEC, x0, x1, v)
        np.testing.assert_allclose(y, want, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(y1, want[:, 0], rtol=1e-13, atol=1e-13)
        # every rank must hold the same full vector (a solver iterates on it)
        t = torch.from_numpy(np.ascontiguousarray(y))
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        # Gram matrix: each rank owns its row block, no collective
        Kb = op.gram_block()
        np.testing.assert_array_equal(Kb, orc.gram(SPEC, x0[op.lo:op.hi], x1))
        # symmetric operator, one column: partial vectors + all-reduce
        ops = ShardedCovarianceOperator(k, x0, None, local_factory=OracleLocal)
        assert ops.local_sym is not None
        ys = ops @ v[:n0, 0] if n0 <= n1 else ops @ np.resize(v[:, 0], n0)
        vs = v[:n0, 0] if n0 <= n1 else np.resize(v[:, 0], n0)
        np.testing.assert_allclose(ys, orc.matvec(SPEC, x0, None, vs[:, None])[:, 0], rtol=1e-12, atol=1e-13)
        ys2 = ops @ np.stack([vs, vs], axis=1)  # two columns: row-sharded path
        np.testing.assert_allclose(ys2[:, 1], ys, rtol=1e-12, atol=1e-13)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n0,n1", [(101, 64), (64, 33)])
def test_sharded_operator_world_size_2(n0, n1):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n0, n1, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert list(out) == [1] * world


# ---------------------------------------------------------------------------------------------
# the reference's own caller over two ranks: problinsolve on k.linop(X) + sigma^2 I with the drop-in
# installed in sharded mode (BASELINE.json configs[3]); per-rank compute is the oracle stand-in
# ---------------------------------------------------------------------------------------------
def _solver_worker(rank, world, port, n, out):
    import sys
    import warnings

    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    sys.path.insert(0, ref)
    import probnum_b200.patch as pb_patch

    pb_patch.numpy2_compat()
    warnings.simplefilter("ignore")
    import probnum as pn

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        X = rng.normal(size=(n, 3))
        y = rng.normal(size=n)
        k = pn.randprocs.covfuncs.Matern((3,), nu=1.5, lengthscales=0.8)
        A_ref = k.linop(X) + 1e-2 * pn.linops.Identity(n)          # reference operator (dense cached)
        x_ref, _, _, info_ref = pn.linalg.problinsolve(A_ref, y, maxiter=15)
        pb_patch.enable(_local_factory=OracleLocal)                  # sharded: a process group with 2 ranks exists
        try:
            L = k.linop(X)
            assert type(L).__name__ == "GpuShardedCovarianceLinearOperator"
            x_gpu, _, _, info = pn.linalg.problinsolve(L + 1e-2 * pn.linops.Identity(n), y, maxiter=15)
        finally:
            pb_patch.disable()
        assert info["iter"] == info_ref["iter"]
        np.testing.assert_allclose(x_gpu.mean, x_ref.mean, rtol=1e-8, atol=1e-10)
        t = torch.from_numpy(np.ascontiguousarray(x_gpu.mean))
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        assert all(torch.equal(g, gathered[0]) for g in gathered)  # replicas stay in lock step
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not os.path.isdir(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                   "baseline", "_ref", "probnum")),
                    reason="needs the installed reference (baseline/_ref)")
def test_problinsolve_over_two_ranks_with_the_sharded_dropin():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_solver_worker, args=(r, world, port, 150, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert list(out) == [1] * world

"""Row-sharded covariance operator: one process per GPU over ``torch.distributed``.

K(X0, X1) is partitioned by contiguous row blocks of X0: rank g keeps
rows ``[lo_g, hi_g)`` of X0 and all of X1 resident on its GPU, computes its slice of
``Y = K @ V`` with the matrix-free kernel and the slices are reassembled with ONE
all-gather (NCCL over NVLink for device tensors, gloo in the CPU tests).  The Gram
matrix shards the same way and needs no collective at all: each rank owns its row
block.  The operand ``V`` is replicated (every rank of a linear solver holds the same
iterate), so nothing else is exchanged.

This is the only communication step of the path; it moves N0*K*8 bytes per product
(8 MiB at N = 2^20, K = 1) next to O(N0*N1*d) flops, so it is latency-sized.

Symmetric operators (``x1 is None``) with one right-hand side take the reduce form of the same
exchange ("symmetric-halving ... reduce-scatter"): every rank holds the whole point
set, evaluates each off-diagonal tile of its round-robin share of row blocks ONCE for both
``y_I`` and ``y_J`` (``pncov_matmul_sym_part_device``), and the full-length partial vectors are
summed with one all-reduce.  Half the kernel evaluations per product; still one collective of
N0*8 bytes.
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from . import _lib


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row block of rank ``rank``: sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("invalid rank / world size")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist  # noqa: PLC0415

    return dist


class ShardedCovarianceOperator:
    """``K(x0, x1)`` with the rows of ``x0`` split over the ranks of ``group``.

    ``local_factory(params, x0_block, x1)`` builds the per-rank operator; it defaults to
    the CUDA operator and exists so that the CPU (gloo) tests can exercise the sharding
    and gather logic with a stand-in.  The object exposes the same ``shape`` / ``dtype`` /
    ``@`` surface as :class:`probnum_b200.covfuncs.CovarianceLinearOperator`.
    """

    def __init__(self, covfunc, x0, x1=None, *, group=None,
                 local_factory: Optional[Callable] = None, device: Optional[int] = None,
                 symmetric: bool = True):
        dist = _dist()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        x0 = covfunc._preprocess_linop_input(x0, argnum=0)
        is_symmetric = x1 is None
        x1 = x0 if x1 is None else covfunc._preprocess_linop_input(x1, argnum=1)
        self.shape = (x0.shape[0], x1.shape[0])
        dt = covfunc._compute_dtype(x0, x1)
        self.dtype = dt
        self.lo, self.hi = shard_bounds(self.shape[0], self.world, self.rank)
        self.max_rows = shard_bounds(self.shape[0], self.world, 0)[1]
        rows = covfunc._points(x0[self.lo:self.hi], dt)
        cols = covfunc._points(x1, dt)
        if local_factory is None:
            self.local = _lib.Operator(covfunc._kernel_params(), rows, cols, device=device)
        else:
            self.local = local_factory(covfunc._kernel_params(), rows, cols)
        # symmetric halving: a second operator over the whole point set (x1 = None).  Whether the
        # symmetric kernel applies to this data is known after the first product (the library
        # answers NotImplementedError, identically on every rank); then the row shards serve.
        self.local_sym = None
        params = covfunc._kernel_params()
        if symmetric and is_symmetric and self.world > 1 and dt == np.float64:
            if local_factory is None:
                if params.matrix_free and params.d <= 96 and self.shape[0] >= 2048:
                    self.local_sym = _lib.Operator(params, cols, None, device=device)
            else:
                self.local_sym = local_factory(params, cols, None)

    # -- host operands (NumPy in, NumPy out; gloo or NCCL via a staging tensor) -------------------
    def matmul(self, v: np.ndarray) -> np.ndarray:
        """Full ``K @ v`` on every rank; ``v`` is (N1,) or (N1, K) and replicated."""
        import torch  # noqa: PLC0415

        dist = _dist()
        vec = v.ndim == 1
        v2 = np.ascontiguousarray(v.reshape(v.shape[0], -1), dtype=self.dtype)
        k = v2.shape[1]
        if self.local_sym is not None and (k == 1 or not hasattr(self.local_sym, "matmul_sym_part")):
            y = self._matmul_symmetric_host(v2)
            if y is not None:
                return y[:, 0] if vec else y
        y_local = self.local.matmul(v2)  # (hi - lo, k)
        pad = np.zeros((self.max_rows, k), dtype=self.dtype)
        pad[: self.hi - self.lo] = y_local
        backend = dist.get_backend(self.group)
        send = torch.from_numpy(pad)
        if backend == "nccl":
            send = send.cuda(self.local.device)
        gathered = torch.empty((self.world * self.max_rows, k), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(gathered, send, group=self.group)
        g = gathered.cpu().numpy().reshape(self.world, self.max_rows, k)
        sizes = [hi - lo for lo, hi in (shard_bounds(self.shape[0], self.world, r) for r in range(self.world))]
        out = np.concatenate([g[r, :sizes[r]] for r in range(self.world)], axis=0)
        return out[:, 0] if vec else out

    def _matmul_symmetric_host(self, v2: np.ndarray) -> Optional[np.ndarray]:
        """One column, symmetric operator: this rank's row blocks + all-reduce."""
        import torch  # noqa: PLC0415

        dist = _dist()
        backend = dist.get_backend(self.group)
        if hasattr(self.local_sym, "matmul_sym_part"):  # CPU stand-in of the gloo tests
            part = torch.from_numpy(np.ascontiguousarray(self.local_sym.matmul_sym_part(v2, self.rank, self.world)))
            if backend == "nccl":
                part = part.cuda()
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
            return part.cpu().numpy()
        vd = torch.from_numpy(v2).cuda(self.local_sym.device)
        y = self._sym_product_device(vd)
        return None if y is None else y.cpu().numpy()

    def _sym_product_device(self, v, out=None):
        """Full product (CUDA tensor (N, K), on every rank) through the symmetric-halving kernels, or None if they
        do not apply.  Every rank accumulates its row blocks into integer planes (pncov.h:
        pncov_matmulk_sym_part_bins_device; passes of up to eight columns), the planes are summed by ONE all-reduce
        each (int64 for the two integer planes, float64 for the plane of non-finite contributions) and converted:
        integer addition is associative, so the result has the same bits for every number of ranks."""
        import torch  # noqa: PLC0415

        dist = _dist()
        n, k = self.shape[0], v.shape[1]
        stream = torch.cuda.current_stream(v.device).cuda_stream
        if out is None:
            out = torch.empty((n, k), dtype=torch.float64, device=v.device)
        for c0 in range(0, k, 8):
            kk = min(8, k - c0)
            words = n if kk == 1 else 8 * n  # per plane
            bins = torch.empty(3 * words, dtype=torch.int64, device=v.device)
            try:
                self.local_sym.matmulk_sym_part_bins_device(v.data_ptr() + 8 * c0, kk, k, bins.data_ptr(), self.rank, self.world,
                                                            stream=stream)
            except NotImplementedError:
                if k == 1:  # no symmetric kernel for this data at all
                    self.local_sym.close()
                    self.local_sym = None
                return None
            dist.all_reduce(bins[: 2 * words], op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(bins[2 * words:].view(torch.float64), op=dist.ReduceOp.SUM, group=self.group)
            self.local_sym.matmulk_sym_finish_device(bins.data_ptr(), kk, out.data_ptr() + 8 * c0, k, stream=stream)
        return out

    def __matmul__(self, v):
        return self.matmul(np.asarray(v))

    # -- Gram matrix: row blocks per rank, no collective --------------------------------
    def gram_block(self) -> np.ndarray:
        """This rank's rows ``[lo, hi)`` of the Gram matrix, shape (hi - lo, N1), as a host array
        (pinned when large).  Stacking the blocks of all ranks in rank order gives ``K(x0, x1)``."""
        from . import hostmem  # noqa: PLC0415

        rows, cols = self.hi - self.lo, self.shape[1]
        out = hostmem.empty((rows, cols), getattr(self.local, "gram_dtype", self.dtype))
        if rows and cols:
            self.local.gram(out)
        return out

    # -- device operands (torch CUDA tensors; NCCL all-gather on the current stream) ----------------
    def matmul_device(self, v, out=None):
        """``v``: CUDA tensor (N1, K) replicated on every rank.  Returns the full
        (N0, K) product on every rank.  Equal shards are required (N0 % world == 0) so the
        all-gather writes straight into the result."""
        import torch  # noqa: PLC0415

        dist = _dist()
        n0, n1 = self.shape
        if n0 % self.world:
            raise ValueError("matmul_device needs N0 divisible by the number of ranks")
        if v.dim() != 2 or v.shape[0] != n1 or not v.is_contiguous():
            raise ValueError("v must be a contiguous (N1, K) CUDA tensor")
        k = v.shape[1]
        rows = self.hi - self.lo
        if self.local_sym is not None and v.dtype == torch.float64:
            if out is not None and not (out.is_contiguous() and out.dtype == torch.float64 and tuple(out.shape) == (n0, k)):
                raise ValueError("out must be a contiguous float64 (N0, K) CUDA tensor")
            y = self._sym_product_device(v, out)
            if y is not None:
                return y
        if out is None:
            out = torch.empty((n0, k), dtype=v.dtype, device=v.device)
        y_local = out[self.lo:self.hi]  # contiguous row block of the result
        stream = torch.cuda.current_stream(v.device).cuda_stream
        self.local.matmul_device(v.data_ptr(), k, k, y_local.data_ptr(), k, stream=stream)
        if self.world > 1:
            dist.all_gather_into_tensor(out, y_local, group=self.group)
        assert rows * self.world == n0
        return out

    def close(self):
        if hasattr(self.local, "close"):
            self.local.close()
        if self.local_sym is not None and hasattr(self.local_sym, "close"):
            self.local_sym.close()
            self.local_sym = None

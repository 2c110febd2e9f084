"""Hand-written FP64 Cholesky (csrc/pncov_chol.cu) next to cuSOLVER potrf (through torch) on one B200.

    python tools/bench_cholesky.py [n ...] > gpurun_out/cholesky.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnum_b200 import _lib  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]
out = {}
for n in sizes:
    torch.manual_seed(n)
    B = torch.randn((n, n + 64), dtype=torch.float64, device="cuda")
    A = B @ B.T / n
    A.diagonal().add_(0.1)
    del B
    flops = n ** 3 / 3.0
    # ours: device input, device time of the factorisation alone (CUDA events inside the library)
    ch = _lib.Cholesky(A.data_ptr(), on_device=True, n=n)
    ch.close()
    t0 = time.perf_counter()
    ch = _lib.Cholesky(A.data_ptr(), on_device=True, n=n)
    wall = time.perf_counter() - t0
    ms = ch.last_factor_ms()
    b = np.random.default_rng(0).normal(size=(n, 2))
    ch.solve(b)
    t0 = time.perf_counter()
    x = ch.solve(b)
    solve_s = time.perf_counter() - t0
    solve_ms = {}
    for k in (1, 2, 3, 6, 12, 16):  # wall time of pncov_chol_solve with host operands, best of 5
        bk = np.random.default_rng(k).normal(size=(n, k))
        ch.solve(bk)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            ch.solve(bk)
            ts.append(time.perf_counter() - t0)
        solve_ms[str(k)] = 1e3 * min(ts)
    # cuSOLVER through torch
    torch.linalg.cholesky_ex(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Lt, info = torch.linalg.cholesky_ex(A)
    e1.record()
    torch.cuda.synchronize()
    ms_lib = e0.elapsed_time(e1)
    xb = torch.cholesky_solve(torch.from_numpy(b).cuda(), Lt).cpu().numpy()
    resid = float(np.linalg.norm((A @ torch.from_numpy(x).cuda()).cpu().numpy() - b) / np.linalg.norm(b))
    out[str(n)] = {
        "ours_factor_ms": ms, "ours_tflops": flops / (ms * 1e-3) / 1e12, "ours_wall_with_copy_s": wall,
        "ours_solve_2rhs_s": solve_s, "ours_solve_ms_by_columns": solve_ms, "cusolver_factor_ms": ms_lib, "cusolver_tflops": flops / (ms_lib * 1e-3) / 1e12,
        "solution_rel_diff_vs_cusolver": float(np.abs(x - xb).max() / np.abs(xb).max()), "rel_residual": resid,
    }
    ch.close()
    del A, Lt
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))

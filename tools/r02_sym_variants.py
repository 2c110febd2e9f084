"""Experiment: variants of the bulk-store symmetric Gram kernel (Matern-5/2, N = 32768, d = 3).
Each variant runs in a fresh process (the variant is read once per process from the environment)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import probnum_b200 as pb
    from probnum_b200 import _lib

    n = int(sys.argv[2])
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev).cuda_stream
    x = torch.from_numpy(np.random.default_rng(42).normal(size=(n, 3))).to(dev)
    out = torch.empty((n, n), dtype=torch.float64, device=dev)
    k = pb.Matern((3,), nu=2.5)
    op = _lib.Operator(k._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
    for _ in range(3):
        op.gram_device(out.data_ptr(), n, stream=stream)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op.gram_device(out.data_ptr(), n, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # checksum for bit-identity across variants + symmetry + unit diagonal
    chk = float(out.sum().item())
    sym = bool(torch.equal(out, out.T))
    diag = bool((out.diagonal() == 1.0).all().item())
    h = int(out.view(torch.int64).sum().item())
    print(f"variant={os.environ.get('PNCOV_GRAM_SYM_VARIANT','0')} tpc={os.environ.get('PNCOV_GRAM_SYM_TPC','16')} "
          f"ms={ms:.4f} GB/s={8.0*n*n/ms/1e6:.0f} sum={chk:.17g} bits={h} sym={sym} diag1={diag}", flush=True)
    sys.exit(0)

n = 32768
for variant, tpcs in ((0, (16,)), (3, (8, 16)), (8, (8, 16, 32)), (9, (16,)), (10, (8, 16, 32)), (11, (16,)), (12, (8, 16, 32))):
    for tpc in tpcs:
        env = dict(os.environ, PNCOV_GRAM_SYM_VARIANT=str(variant), PNCOV_GRAM_SYM_TPC=str(tpc))
        r = subprocess.run([sys.executable, __file__, "child", str(n)], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or ("FAILED: " + r.stderr[-400:]), flush=True)

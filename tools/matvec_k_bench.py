"""Device-timed ExpQuad K @ V for several numbers of right-hand sides (d = 8)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb
from probnum_b200 import _lib
from oracle import covfuncs_oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
d = 8
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
rng = np.random.default_rng(42)
xh = rng.normal(size=(n, d))
x = torch.from_numpy(xh).to(dev)
op = _lib.Operator(pb.ExpQuad((d,))._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
for k in (1, 2, 4, 5, 8, 16):
    vh = rng.uniform(size=(n, k))
    v = torch.from_numpy(vh).to(dev)
    y = torch.empty((n, k), dtype=torch.float64, device=dev)
    op.matmul_device(v.data_ptr(), k, k, y.data_ptr(), k, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    op.matmul_device(v.data_ptr(), k, k, y.data_ptr(), k, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rows = np.arange(0, n, n // 16)
    spec = orc.KernelSpec("expquad", (d,), 1.0)
    Kr = orc.gram(spec, xh[rows], xh, row_block=16)
    ratio = np.abs(y[torch.from_numpy(rows).to(dev)].cpu().numpy() - Kr @ vh) / (1e-12 * (np.abs(Kr) @ np.abs(vh)))
    print(f"n={n} K={k} path={op.matmul_path()} ms={ms:.2f} ms_per_column={ms/k:.2f} Gpair_columns/s={n*n*k/ms/1e6:.0f} parity_max_ratio={ratio.max():.2e}", flush=True)
op.close()
# cross operator (x1 != x0): the ordered kernels
x1h = rng.normal(size=(n, d))
x1 = torch.from_numpy(x1h).to(dev)
opx = _lib.Operator(pb.ExpQuad((d,))._kernel_params(), x.data_ptr(), x1.data_ptr(), on_device=True, n0=n, n1=n, dtype=np.float64)
for k in (1, 4, 8):
    vh = rng.uniform(size=(n, k))
    v = torch.from_numpy(vh).to(dev)
    y = torch.empty((n, k), dtype=torch.float64, device=dev)
    opx.matmul_device(v.data_ptr(), k, k, y.data_ptr(), k, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    opx.matmul_device(v.data_ptr(), k, k, y.data_ptr(), k, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rows = np.arange(0, n, n // 16)
    Kr = orc.gram(orc.KernelSpec("expquad", (d,), 1.0), xh[rows], x1h, row_block=16)
    ratio = np.abs(y[torch.from_numpy(rows).to(dev)].cpu().numpy() - Kr @ vh) / (1e-12 * (np.abs(Kr) @ np.abs(vh)))
    print(f"cross n={n} K={k} path={opx.matmul_path()} ms={ms:.2f} ms_per_column={ms/k:.2f} parity_max_ratio={ratio.max():.2e}", flush=True)
opx.close()

import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import probnum_b200 as pb
from probnum_b200 import _lib
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(dev).cuda_stream
for d in (2, 3, 4, 5, 6, 8):
    n = 262144
    x = torch.from_numpy(np.random.default_rng(0).normal(size=(n, d))).to(dev)
    v = torch.randn((n, 1), dtype=torch.float64, device=dev); y = torch.empty_like(v)
    res = []
    for form in ("dmma", "dfma", "off"):
        os.environ["PNCOV_MATVEC_IDENT"] = form
        op = _lib.Operator(pb.ExpQuad((d,), lengthscales=np.sqrt(d))._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
        for _ in range(3): op.matmul_device(v.data_ptr(), 1, 1, y.data_ptr(), 1, stream=stream)
        torch.cuda.synchronize()
        res.append("%s %.1f" % (form, n * n / op.last_kernel_ms() / 1e6))
        op.close()
    print("d=%d Gpair/s:" % d, "  ".join(res))

"""Device-timed dense Gram at wide d (ExpQuad / RatQuad): guarded norm-identity DMMA kernel vs the
direct-difference chunk kernel (PNCOV_GRAM_IDENT=off), N = 16384 by default."""
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
    for d in [int(a) for a in sys.argv[3:]]:
        x = torch.from_numpy(np.random.default_rng(42).normal(size=(n, d))).to(dev)
        x1 = torch.from_numpy(np.random.default_rng(43).normal(size=(n, d))).to(dev)
        out = torch.empty((n, n), dtype=torch.float64, device=dev)
        for name, k in (("expquad", pb.ExpQuad((d,), lengthscales=0.7 * d ** 0.5)), ("ratquad", pb.RatQuad((d,), lengthscale=0.7 * d ** 0.5, alpha=1.0)),
                        ("ratquad_a0.7", pb.RatQuad((d,), lengthscale=0.7 * d ** 0.5, alpha=0.7))):
            for label, other in (("sym", None), ("cross", x1)):
                op = _lib.Operator(k._kernel_params(), x.data_ptr(), None if other is None else other.data_ptr(), on_device=True, n0=n, n1=n, dtype=np.float64)
                for _ in range(2):
                    op.gram_device(out.data_ptr(), n, stream=stream)
                torch.cuda.synchronize()
                reps = 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    op.gram_device(out.data_ptr(), n, stream=stream)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print(f"ident={os.environ.get('PNCOV_GRAM_IDENT','on')} {name} {label} n={n} d={d} ms={ms:.4f} "
                      f"GB/s={8.0*n*n/ms/1e6:.0f} dmma_TFLOPs={2.0*n*n*d/ms/1e9:.1f} sum={float(out.sum().item()):.15g}", flush=True)
                op.close()
    sys.exit(0)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ds = sys.argv[2:] or ["17", "32", "64", "100", "256"]
for ident in ("on", "off"):
    env = dict(os.environ, PNCOV_GRAM_IDENT=ident)
    r = subprocess.run([sys.executable, __file__, "child", str(n)] + ds, env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout.strip() or ("FAILED: " + r.stderr[-800:]), flush=True)

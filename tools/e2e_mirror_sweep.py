"""End-to-end Matern-5/2 Gram (N = 32768, d = 3) through pncov_gram into a pinned host array for several
settings of the symmetric return path (each in a fresh process: the settings are read once)."""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import probnum_b200 as pb
    from probnum_b200 import _lib, hostmem
    n = 32768
    x = np.random.default_rng(42).normal(size=(n, 3))
    K = hostmem.empty((n, n), np.float64)
    with _lib.Operator(pb.Matern((3,), nu=2.5)._kernel_params(), x, None) as op:
        op.gram(K)
        ts = []
        for _ in range(6):
            t0 = time.perf_counter(); op.gram(K); ts.append(time.perf_counter() - t0)
    print(f"mirror={os.environ.get('PNCOV_GRAM_MIRROR','on')} block_mib={os.environ.get('PNCOV_GRAM_MIRROR_BLOCK_MIB','256')} "
          f"full_div={os.environ.get('PNCOV_GRAM_MIRROR_FULL_DIV','16')} threads={os.environ.get('PNCOV_HOST_THREADS','all')} "
          f"ms min={1e3*min(ts):.1f} median={1e3*sorted(ts)[len(ts)//2]:.1f} sym={bool(np.array_equal(K[:2048, 30000:], K[30000:, :2048].T))}", flush=True)
    sys.exit(0)
configs = [dict(PNCOV_GRAM_MIRROR="off")]
for mib in ("1024", "256", "64"):
    for div in ("0", "16", "8", "4"):
        configs.append(dict(PNCOV_GRAM_MIRROR_BLOCK_MIB=mib, PNCOV_GRAM_MIRROR_FULL_DIV=div))
configs.append(dict(PNCOV_GRAM_MIRROR_BLOCK_MIB="1024", PNCOV_GRAM_MIRROR_FULL_DIV="0", PNCOV_HOST_THREADS="8"))
configs.append(dict(PNCOV_GRAM_MIRROR="off"))
for cfg in configs:
    r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **cfg), capture_output=True, text=True, timeout=600)
    print(r.stdout.strip() or ("FAILED " + r.stderr[-500:]), flush=True)

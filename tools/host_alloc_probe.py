"""Page-locked result blocks: allocation time and end-to-end symmetric Gram time for the two allocators
(PNCOV_HOST_ALLOC=cuda: cudaHostAlloc; default: mmap + huge-page advice + parallel touch + cudaHostRegister)."""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import probnum_b200 as pb
    from probnum_b200 import _lib, hostmem
    n = 32768
    x = np.random.default_rng(42).normal(size=(n, 3))
    _lib.load()
    import torch; torch.cuda.init(); torch.zeros(1).cuda()
    t0 = time.perf_counter(); K = hostmem.empty((n, n), np.float64); t_alloc = time.perf_counter() - t0
    with _lib.Operator(pb.Matern((3,), nu=2.5)._kernel_params(), x, None) as op:
        res = {}
        for mode in ("on", "off"):
            os.environ["PNCOV_GRAM_MIRROR"] = mode
        # env is read per handle: two handles
    for mode in ("on", "off"):
        os.environ["PNCOV_GRAM_MIRROR"] = mode
        with _lib.Operator(pb.Matern((3,), nu=2.5)._kernel_params(), x, None) as op:
            op.gram(K)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); op.gram(K); ts.append(time.perf_counter() - t0)
            res[mode] = 1e3 * min(ts)
    thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists("/sys/kernel/mm/transparent_hugepage/enabled") else "?"
    huge = [l for l in open("/proc/self/smaps_rollup") if "AnonHugePages" in l]
    print(f"alloc={os.environ.get('PNCOV_HOST_ALLOC','mmap')} alloc_8.6GB_s={t_alloc:.2f} gram_mirror_ms={res['on']:.1f} gram_full_ms={res['off']:.1f} thp={thp} {huge[0].strip() if huge else ''}", flush=True)
    sys.exit(0)
for alloc in ("mmap4k", "mmap", "cuda", "mmap4k", "mmap"):
    r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, PNCOV_HOST_ALLOC=alloc), capture_output=True, text=True, timeout=600)
    print(r.stdout.strip() or ("FAILED " + r.stderr[-600:]), flush=True)

"""Host-side probe of a GPU box: cores, NUMA, memory bandwidth with 1..all threads (memcpy
through numpy, which releases the GIL), used to size the host-side mirror of symmetric Grams."""
import os, time, threading, subprocess
import numpy as np

print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print(subprocess.run(["lscpu"], capture_output=True, text=True).stdout)
    print(subprocess.run(["free", "-g"], capture_output=True, text=True).stdout)
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
except OSError as e:
    print("probe failed", e)

n = 1 << 27  # 1 GiB of float64 per thread pair
for threads in (1, 2, 4, 8, 16, 32):
    if threads > (os.cpu_count() or 1):
        break
    srcs = [np.ones(n // threads) for _ in range(threads)]
    dsts = [np.empty(n // threads) for _ in range(threads)]
    def work(i):
        for _ in range(4):
            np.copyto(dsts[i], srcs[i])
    for rep in range(2):
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in ts]; [t.join() for t in ts]
        dt = time.perf_counter() - t0
    print(f"threads {threads:2d}: copy {4 * 8 * n / dt / 1e9:.1f} GB/s (read+write counted once each: {2 * 4 * 8 * n / dt / 1e9:.1f})")

"""Prediction path: python tools/bench_predict.py [n] [m] [d]

Times pncov_chol_solve_gram -- cross Gram K(nodes, x), W = (L L^T)^-1 K, column sums of W * K and
vecs @ W in one device pass -- against (a) the same work through the generic pncov_chol_solve
(host in / host out, 8 columns per sweep), (b) cuSOLVER potrs through torch on a device-resident
Gram block, (c) scipy.linalg.cho_solve on the host for a bounded column sample.  One JSON line."""
import json
import os
import sys
import time

import numpy as np
import scipy.linalg
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
d = int(sys.argv[3]) if len(sys.argv) > 3 else 4
device_only = len(sys.argv) > 4 and sys.argv[4] == "device-only"  # for runs under ncu
rng = np.random.default_rng(42)
nodes, x = rng.normal(size=(n, d)), rng.normal(size=(m, d))
fun_evals = np.sin(nodes).sum(axis=1)
k = pb.ExpQuad((d,))
jitter = 1e-8
res = {"workload": f"ExpQuad d={d}, {n} nodes, {m} prediction points, jitter {jitter}, float64"}

with _lib.Operator(k._kernel_params(), nodes) as sq:
    ch = _lib.Cholesky(sq, jitter=jitter)
res["factor_ms"] = ch.last_factor_ms()
with _lib.Operator(k._kernel_params(), nodes, x) as op:
    for _ in range(2):
        t0 = time.perf_counter()
        quad, dots = ch.solve_gram(op, [fun_evals])
        wall = time.perf_counter() - t0
    res["solve_gram_device_ms"] = ch.last_solve_gram_ms
    res["solve_gram_wall_ms"] = wall * 1e3
    flops = 2.0 * n * n * m  # forward + backward substitution
    res["solve_gram_tflops"] = flops / (ch.last_solve_gram_ms * 1e-3) / 1e12
    for _ in range(2):
        quad_f, dots_f = ch.solve_gram(op, [fun_evals], forward_only=True)
    res["forward_only_device_ms"] = ch.last_solve_gram_ms
    res["forward_only_tflops"] = 0.5 * flops / (ch.last_solve_gram_ms * 1e-3) / 1e12
    res["forward_only_vs_reference_order_abs"] = [float(np.abs(quad_f - quad).max()), float(np.abs(dots_f - dots).max())]
if device_only:
    print(json.dumps(res))
    sys.exit(0)

# (a) generic solve on a column sample
ms_ = min(m, 64)
G = k.matrix(nodes, x[:ms_])
t0 = time.perf_counter()
W = ch.solve(G)
res["generic_solve_ms_per_column"] = (time.perf_counter() - t0) * 1e3 / ms_
err = np.abs(np.sum(W * G, axis=0) - quad[:ms_]).max()
res["quad_vs_generic_solve_abs"] = float(err)
res["dots_vs_generic_solve_abs"] = float(np.abs(fun_evals @ W - dots[0, :ms_]).max())

# (b) cuSOLVER potrs through torch
dev = torch.device("cuda", 0)
L = torch.from_numpy(ch.factor()).to(dev)
Gd = torch.from_numpy(k.matrix(nodes, x)).to(dev)
for _ in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Wd = torch.cholesky_solve(Gd, L, upper=False)
    q = (Wd * Gd).sum(dim=0)
    e1.record()
    torch.cuda.synchronize()
res["cusolver_potrs_ms"] = e0.elapsed_time(e1)
res["quad_vs_cusolver_abs"] = float(np.abs(q.cpu().numpy() - quad).max())
del L, Gd, Wd

# (c) host scipy on a bounded sample (factor once, untimed; the solve is what is compared)
if n <= 16384:
    K = k.matrix(nodes)
    K[np.diag_indices(n)] += jitter
    cf_host = scipy.linalg.cho_factor(K, lower=True)
    t0 = time.perf_counter()
    Wh = scipy.linalg.cho_solve(cf_host, G)
    res["scipy_cho_solve_ms_per_column"] = (time.perf_counter() - t0) * 1e3 / ms_
    res["host_threads"] = os.cpu_count()
print(json.dumps(res))

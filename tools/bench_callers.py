"""Caller-level runs of BASELINE.json configs[0], [3] and [4] with the device path installed
into the real ProbNum (baseline/_ref), timed next to the unmodified reference where the
reference can run at all.

    python tools/bench_callers.py [--n-solve 65536] [--n-bq 16384] [--maxiter 50] > profiles/r01_callers.json

configs[0]  ExpQuad Gram N=1000, d=10 (and the asv shape d=100, benchmarks/covfuncs.py:57)
configs[3]  problinsolve on (K_ExpQuad + 1e-3 I) x = y, N=65536, d=4, matrix-free GPU operator
configs[4]  bayesquad_from_data, ExpQuad + Gaussian measure, 16384 nodes, d=4
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import probnum_b200.patch as gpu  # noqa: E402

gpu.numpy2_compat()
warnings.simplefilter("ignore")
import probnum as pn  # noqa: E402

from probnum_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-solve", type=int, default=65536)
ap.add_argument("--n-solve-cpu", type=int, default=8192)
ap.add_argument("--n-bq", type=int, default=16384)
ap.add_argument("--n-bq-cpu", type=int, default=4096)
ap.add_argument("--maxiter", type=int, default=50)
args = ap.parse_args()
cf = pn.randprocs.covfuncs
out = {"cpu_count": os.cpu_count()}


def best_of(fn, reps=3):
    best, val = 1e30, None
    for _ in range(reps):
        t0 = time.perf_counter()
        val = fn()
        best = min(best, time.perf_counter() - t0)
    return best, val


# ---- configs[0]: the reference's own asv shapes ------------------------------------------------
for d in (10, 100):
    X = np.random.default_rng(42).normal(size=(1000, d))
    k = cf.ExpQuad((d,))
    t_cpu, K_cpu = best_of(lambda: k.matrix(X))
    gpu.enable()
    k.matrix(X)  # warm-up (context, pinned pool)
    t_gpu, K_gpu = best_of(lambda: k.matrix(X), reps=10)
    gpu.disable()
    out[f"config0_expquad_gram_n1000_d{d}"] = {
        "cpu_ms": 1e3 * t_cpu, "gpu_api_ms": 1e3 * t_gpu, "speedup": t_cpu / t_gpu,
        "max_abs_diff": float(np.abs(K_gpu - K_cpu).max()),
    }

# ---- configs[3]: problinsolve ------------------------------------------------------------------
def solve(n, ell, maxiter, timing):
    rng = np.random.default_rng(42)
    X = rng.normal(size=(n, 4))
    y = rng.normal(size=n)
    k = cf.ExpQuad((4,), lengthscales=ell)
    t0 = time.perf_counter()
    L = k.linop(X)
    A = L + 1e-3 * pn.linops.Identity(n)
    x, _, _, info = pn.linalg.problinsolve(A, y, maxiter=maxiter)
    timing["total_s"] = time.perf_counter() - t0
    timing["iter"] = int(info["iter"])
    timing["resid_l2norm"] = float(info["resid_l2norm"])
    return x.mean


for ell in (1.0, 0.25):
    tim_cpu, tim_gpu, tim_big = {}, {}, {}
    x_cpu = solve(args.n_solve_cpu, ell, args.maxiter, tim_cpu)  # dense-cached reference operator
    gpu.enable()
    launches0 = _lib.launch_count()
    x_gpu = solve(args.n_solve_cpu, ell, args.maxiter, tim_gpu)
    x_big = solve(args.n_solve, ell, args.maxiter, tim_big)
    tim_big["kernel_launches"] = _lib.launch_count() - launches0
    gpu.disable()
    out[f"config3_problinsolve_ell{ell}"] = {
        "n_compare": args.n_solve_cpu, "reference_cpu": tim_cpu, "gpu_operator": tim_gpu,
        "x_rel_diff_inf": float(np.abs(x_gpu - x_cpu).max() / np.abs(x_cpu).max()),
        "n_full": args.n_solve, "gpu_operator_full": tim_big,
        "note": "the reference cannot run n_full: its operator caches the dense matrix (34 GB at 65536)",
    }

# ---- configs[4]: Bayesian quadrature -------------------------------------------------------------
def bq(n, timing):
    rng = np.random.default_rng(42)
    nodes = rng.normal(size=(n, 4))
    f = np.sin(nodes).sum(axis=1)
    measure = pn.quad.integration_measures.GaussianMeasure(mean=np.zeros(4), cov=np.eye(4))
    t0 = time.perf_counter()
    F, _ = pn.quad.bayesquad_from_data(nodes=nodes, fun_evals=f, measure=measure)
    timing["total_s"] = time.perf_counter() - t0
    return float(F.mean), float(F.var)


tim_cpu, tim_gpu, tim_big = {}, {}, {}
m_cpu, v_cpu = bq(args.n_bq_cpu, tim_cpu)
gpu.enable()
m_gpu, v_gpu = bq(args.n_bq_cpu, tim_gpu)
m_big, v_big = bq(args.n_bq, tim_big)
t0 = time.perf_counter()
K = cf.ExpQuad((4,)).matrix(np.random.default_rng(42).normal(size=(args.n_bq, 4)))
tim_big["gram_only_s"] = time.perf_counter() - t0
import tools.cusolver_baseline as cusolver_baseline  # noqa: E402

for backend in ("cusolver", "pncov"):
    if backend == "pncov":
        cusolver_baseline.disable()
        gpu.enable_quad_cholesky()
    else:
        cusolver_baseline.enable()
    tim_chol = {}
    m_ch, v_ch = bq(args.n_bq, tim_chol)  # first call pays library initialisation / allocation
    tim_chol2 = {}
    m_ch, v_ch = bq(args.n_bq, tim_chol2)
    out[f"config4_bayesquad_device_cholesky_{backend}"] = {
        "n": args.n_bq, "first_call": tim_chol, "second_call": tim_chol2, "mean": m_ch, "var": v_ch,
        "mean_abs_diff_vs_host_cholesky": abs(m_ch - m_big),
        "note": ("SURVEY 8f N1 baseline: cho_factor / cho_solve through torch (cuSOLVER)" if backend == "cusolver" else
                 "SURVEY 8f N1: hand-written blocked FP64 Cholesky of libpncov (csrc/pncov_chol.cu)")
                + ", Gram still returned to the host",
    }
gpu.disable()
out["config4_bayesquad"] = {
    "n_compare": args.n_bq_cpu, "reference_cpu": tim_cpu, "gpu_gram": tim_gpu,
    "mean_abs_diff": abs(m_gpu - m_cpu), "var_rel_diff": abs(v_gpu - v_cpu) / abs(v_cpu),
    "n_full": args.n_bq, "gpu_gram_full": tim_big, "mean_full": m_big, "var_full": v_big,
    "note": "after the Gram moves to the GPU the host Cholesky (scipy cho_factor) dominates: SURVEY 8f N1",
}
print(json.dumps(out, indent=1))

"""Caller-level run of the prediction path with the device path installed into the
real ProbNum (baseline/_ref): BQStandardBeliefUpdate.predict_integrand and the
IntegralVarianceReduction acquisition on a fitted BQ state, next to the unmodified reference on a
bounded sample of the prediction points.

    python tools/bench_predict_callers.py [n_nodes] [m_points] [m_host] > profiles/r01_predict_callers.json
"""
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
from probnum.quad.solvers import BQState  # noqa: E402
from probnum.quad.solvers.acquisition_functions import IntegralVarianceReduction  # noqa: E402
from probnum.quad.solvers.belief_updates import BQStandardBeliefUpdate  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
m_host = int(sys.argv[3]) if len(sys.argv) > 3 else 128
d = 4
rng = np.random.default_rng(42)
nodes, x = rng.normal(size=(n, d)), rng.normal(size=(m, d))
fun_evals = np.sin(nodes).sum(axis=1)
measure = pn.quad.integration_measures.GaussianMeasure(mean=np.zeros(d), cov=np.eye(d))
cf = pn.randprocs.covfuncs


def fit():
    update = BQStandardBeliefUpdate(jitter=1e-8, scale_estimation="mle")
    state = BQState(measure=measure, kernel=cf.ExpQuad((d,)))
    t0 = time.perf_counter()
    _, state = update(state, nodes, fun_evals)
    return state, time.perf_counter() - t0


def predict(state, pts):
    t0 = time.perf_counter()
    mean, var = BQStandardBeliefUpdate.predict_integrand(pts, state)
    t1 = time.perf_counter()
    acq, _ = IntegralVarianceReduction()(pts, state)
    return mean, var, acq, t1 - t0, time.perf_counter() - t1


out = {"workload": f"ExpQuad d={d}, {n} nodes, jitter 1e-8, mle scale; {m} prediction points on the device, "
                   f"{m_host} of them through the unmodified reference", "host_threads": os.cpu_count()}
state_cpu, fit_cpu = fit()
mean_c, var_c, acq_c, tp_c, ta_c = predict(state_cpu, x[:m_host])
out["reference"] = {"fit_s": fit_cpu, "predict_integrand_s_per_point": tp_c / m_host, "ivr_s_per_point": ta_c / m_host}

gpu.enable()
gpu.enable_quad_cholesky()
for mode in ("reference_order", "forward_only"):
    gpu.PREDICTION_FORWARD_ONLY = mode == "forward_only"
    state, fit_gpu = fit()
    predict(state, x[:256])  # warm-up
    mean_g, var_g, acq_g, tp_g, ta_g = predict(state, x)
    out[mode] = {
        "fit_s": fit_gpu, "predict_integrand_s": tp_g, "ivr_s": ta_g,
        "predict_integrand_s_per_point": tp_g / m, "ivr_s_per_point": ta_g / m,
        "speedup_predict_per_point": (tp_c / m_host) / (tp_g / m), "speedup_ivr_per_point": (ta_c / m_host) / (ta_g / m),
        "mean_max_abs_diff": float(np.abs(mean_g[:m_host] - mean_c).max()),
        "var_max_abs_diff": float(np.abs(var_g[:m_host] - var_c).max()),
        "var_scale": float(np.abs(var_c).max()),
        "acq_max_rel_diff": float(np.max(np.abs(acq_g[:m_host] - acq_c) / np.abs(acq_c))),
    }
gpu.PREDICTION_FORWARD_ONLY = False
gpu.disable()
print(json.dumps(out, indent=1))

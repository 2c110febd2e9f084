import sys, os, time, warnings, cProfile, pstats
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/baseline/_ref')
import numpy as np
import probnum_b200.patch as gpu
gpu.numpy2_compat(); warnings.simplefilter("ignore")
import probnum as pn
from probnum.quad.solvers import BQState
from probnum.quad.solvers.belief_updates import BQStandardBeliefUpdate
d=4
measure = pn.quad.integration_measures.GaussianMeasure(mean=np.zeros(d), cov=np.eye(d))
gpu.enable(); gpu.enable_quad_cholesky()
for n in (8192, 16384):
    rng = np.random.default_rng(42)
    nodes = rng.normal(size=(n, d)); f = np.sin(nodes).sum(axis=1)
    def fit():
        update = BQStandardBeliefUpdate(jitter=1e-8, scale_estimation="mle")
        state = BQState(measure=measure, kernel=pn.randprocs.covfuncs.ExpQuad((d,)))
        t0=time.perf_counter(); _, state = update(state, nodes, f); return time.perf_counter()-t0
    print(n, [round(fit(),4) for _ in range(4)])
    t0=time.perf_counter(); F,_ = pn.quad.bayesquad_from_data(nodes=nodes, fun_evals=f, measure=measure); print("bayesquad_from_data", time.perf_counter()-t0)
    pr=cProfile.Profile(); pr.enable(); fit(); pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(14)

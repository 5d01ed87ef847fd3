"""Host-side fixed cost of PDS1D_Ensemble.solve on C3's shape: the same call with 10 time steps instead of 5000, cProfile on top.
python scratch/e2e_overhead.py"""
import os, sys, time, cProfile, pstats, io
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
w = bench.make_workload("c3")
n = 10
ws = dict(w, n_steps=n, srcs=[(t[: 2] * 0 + np.array([0.0, float(n)]), d[:2]) for t, d in w["srcs"]])
ens = bench.build_ensemble(ws)
obs = np.random.default_rng(1).uniform(0, 1, (n + 1, len(w["obs_idx"])))
ens.set_observations(w["obs_idx"], obs)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ens.solve(dt=1.0, t_total=float(n), audit_models=w["audit"])
    torch.cuda.synchronize(); print("solve (10 steps):", (time.perf_counter() - t0) * 1e3, "ms")
pr = cProfile.Profile(); pr.enable()
for it in range(10):
    ens.solve(dt=1.0, t_total=float(n), audit_models=w["audit"])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:5000])

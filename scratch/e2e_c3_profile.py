"""Where does PDS1D_Ensemble.solve spend host time on C3?  cProfile of one end-to-end call (scratch tool)."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
w = bench.make_workload("c3", 1)
ens = bench.build_ensemble(w)
obs = np.zeros((w["n_steps"] + 1, len(w["obs_idx"])))
ens.set_observations(w["obs_idx"], obs)
for it in range(3):
    ens.solve(dt=1.0, t_total=float(w["n_steps"]), audit_models=w["audit"])
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
ens.solve(dt=1.0, t_total=float(w["n_steps"]), audit_models=w["audit"])
torch.cuda.synchronize()
pr.disable()
print(f"total {1e3 * (time.perf_counter() - t0):.2f} ms")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])

import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fiberis_b200 import _engine as E
w = bench.make_workload("c3", 1)
ens = bench.build_ensemble(w)
obs = np.zeros((w["n_steps"] + 1, 16))
ens.set_observations(w["obs_idx"], obs)
import cProfile, pstats
for i in range(2):
    ens.solve(dt=1.0, t_total=5000.0, audit_models=w["audit"])
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t=time.perf_counter()
ens.solve(dt=1.0, t_total=5000.0, audit_models=w["audit"])
torch.cuda.synchronize()
print("total", time.perf_counter()-t)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
# raw copies
x = torch.empty((5000,4,1024), dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t=time.perf_counter(); y=x.cpu(); print("pageable d2h 164MB", time.perf_counter()-t)
t=time.perf_counter(); pin=torch.empty((5000,4,1024), dtype=torch.float64, pin_memory=True); print("pinned alloc", time.perf_counter()-t)
t=time.perf_counter(); pin.copy_(x, non_blocking=True); torch.cuda.synchronize(); print("pinned d2h", time.perf_counter()-t)
t=time.perf_counter(); z=pin.numpy().copy(); print("host copy", time.perf_counter()-t)

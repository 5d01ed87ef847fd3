"""Ensembles of SHORT meshes (the reference's own model size is 200 cells): cell-updates/s of PDS1D_Ensemble.solve."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_ensemble import make_ensemble
shapes = [(65536, 64, 1000), (65536, 128, 1000), (32768, 200, 1000), (32768, 256, 1000), (16384, 512, 1000)]
if len(sys.argv) > 1:
    shapes = [tuple(int(q) for q in a.split("x")) for a in sys.argv[1:]]
for M, nx, n_steps in shapes:
    e, c = make_ensemble(5, M, nx, n_steps, n_obs=4)
    e.set_observations(c["obs_idx"], np.zeros((n_steps + 1, 4)))
    ts = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e.solve(dt=1.0, t_total=c["T"])
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t = min(ts[1:])
    print(f"M={M:6d} nx={nx:4d} steps={n_steps}: {t*1e3:8.2f} ms  {M*nx*n_steps/t/1e9:8.1f} G cell-updates/s")

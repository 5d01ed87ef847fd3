"""Throughput of the device-side adaptive loop on an ensemble (iterations = 3 solves + 2 assemblies each)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_ensemble import make_ensemble
for M, nx in [(1024, 1024), (4096, 512), (592, 1024)]:
    e, c = make_ensemble(5, M, nx, 200)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = e.solve_adaptive(dt_init=0.05, t_total=200.0, tol=1e-3, max_dt=4.0)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    its = int(out["iterations"].sum()); acc = int(out["accepted"].sum())
    print(f"M={M} nx={nx}: {dt*1e3:.1f} ms, {its} iterations ({acc} accepted), {its/dt/1e6:.2f} M iterations/s, "
          f"{3*its*nx/dt/1e9:.1f} G solve-unknowns/s, {dt/its*M*1e6 if False else dt*1e6/(its/M):.2f} us per iteration per member-slot")

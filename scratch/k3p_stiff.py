"""K3P on stiff long meshes (where it is the path taken): us per step — python scratch/k3p_stiff.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fiberis_b200 import _engine as E
rng = np.random.default_rng(3)
nx = 4_000_000
for dx, label in ((0.03, "alpha ~ 5e2"), (0.004, "alpha ~ 3e4"), (0.0004, "alpha ~ 3e6")):
    mesh = np.cumsum(rng.uniform(0.9 * dx, 1.1 * dx, nx))
    D = np.ones(nx)
    mesh_d, D_d = E.to_device(mesh), E.to_device(D)
    p0 = E.to_device(rng.uniform(0, 1, nx)[None, :])
    sidx, n_src = E.index_tensor(np.array([nx // 3, 2 * nx // 3]))
    n_steps = 40
    table = E.to_device(rng.uniform(1, 10, (n_steps, 2)))
    fac = E.factorize(*E.assemble(mesh_d, D_d, 1, nx, 0.5, "Neumann", "Dirichlet", sidx, n_src), check_singular=False)
    ts = []
    for it in range(4):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = E.run_stream(fac, p0, nx, n_steps, "Neumann", "Dirichlet", sidx, n_src, table, snapshot_every=n_steps,
                           tile_warps=int(os.environ.get("K3P_W", "8")))[0]
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / n_steps)
    print(f"{label}: {min(ts[1:]):.1f} us per step  ({nx * 40.0 / min(ts[1:]) / 1e6:.2f} TB/s at 40 B)  checksum {float(out.sum()):.6e}")

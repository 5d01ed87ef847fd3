"""Single model, every step recorded (the reference's default use): run kernel time per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fiberis_b200 import _engine as E
n_steps = 2000
for nx in (200, 256, 512, 1024, 2048, 4096):
    rng = np.random.default_rng(nx)
    mesh_d = E.to_device(np.cumsum(rng.uniform(0.5, 2.0, nx))); D_d = E.to_device(10 ** rng.uniform(-1, 1, nx))
    sidx_d, n_src = E.index_tensor(np.array([nx // 4, 3 * nx // 4]))
    fac = E.factorize(*E.assemble(mesh_d, D_d, 1, nx, 0.5, "Neumann", "Dirichlet", sidx_d, n_src), check_singular=False)
    p0 = E.to_device(rng.uniform(0, 1, (1, nx))); table_d = E.to_device(rng.uniform(-1, 1, (n_steps, 2)))
    for every in (1, 10):
        ts = []
        for it in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = E.run(fac, p0, 1, nx, n_steps, "Neumann", "Dirichlet", sidx_d, n_src, table_d, snapshot_every=every)
            b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f"{os.environ.get('FBR_K3_GENERIC','x'):>3s} nx={nx:5d} every={every:2d}: {min(ts)*1e3/n_steps:6.3f} us/step")

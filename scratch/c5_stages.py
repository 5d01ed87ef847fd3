"""C5 (1e7 cells): device time of every stage of PDS1D_MultiSource.solve's long-mesh path (scratch tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fiberis_b200 import _engine as E
nx, n_steps, dt = 10_000_000, 100, 0.5
rng = np.random.default_rng(505)
mesh_d = E.to_device(np.cumsum(rng.uniform(0.5, 2.0, nx)))
diff_d = E.to_device(10 ** rng.uniform(-1, 1, nx))
sidx = np.round(np.linspace(.1, .9, 16) * nx).astype(np.int64)
sidx_d, n_src = E.index_tensor(sidx)
p0_d = E.to_device(rng.uniform(0, 1, (1, nx)))
table_d = E.to_device(rng.uniform(-10, 10, (n_steps, 16)))
def timed(label, fn, reps=4):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{label:34s} {min(ts):8.3f} ms")
    return out
diag = timed("K1 assemble", lambda: E.assemble(mesh_d, diff_d, 1, nx, dt, "Neumann", "Dirichlet", sidx_d, n_src))
fac = timed("K1b factorize (cell-parallel)", lambda: E.factorize(*diag, check_singular=False))
timed("K1b factorize + flag check", lambda: E.factorize(*diag))
hor = timed("decay horizon", lambda: E.decay_horizon(fac))
plan = timed("window plan", lambda: E.window_plan(nx, hor, 100))
print(hor, plan)
timed("K3W 100 steps", lambda: E.run_window(fac, p0_d, nx, n_steps, "Neumann", "Dirichlet", sidx_d, n_src, table_d, hor, snapshot_every=100), reps=3)

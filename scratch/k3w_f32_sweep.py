"""fp32 K3W: (warps, steps per block) sweep on C2 / C5 — python scratch/k3w_f32_sweep.py c2|c5 [f32|f64]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fiberis_b200 import _engine as E

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
dt_name = "float32" if (len(sys.argv) < 3 or sys.argv[2] == "f32") else "float64"
w = bench.make_single(wl)
nx, n_steps, dt, every = w["nx"], w["n_steps"], w["dt"], w["every"]
t_axis = np.arange(n_steps) * dt
table = np.ascontiguousarray(np.stack([np.interp(t_axis, st, sd) for st, sd in w["srcs"]], axis=1))
mesh_d, D_d, p0_d = E.to_device(w["mesh"]), E.to_device(w["D"]), E.to_device(w["initial"][None, :])
sidx_d, n_src = E.index_tensor(np.asarray(w["sidx"]))
table_d = E.to_device(table)
fac = E.factorize(*E.assemble(mesh_d, D_d, 1, nx, dt, w["lbc"], w["rbc"], sidx_d, n_src), check_singular=False)
horizon = E.decay_horizon(fac)
print(wl, dt_name, "horizon", horizon)
combos = [(None, None)] + [(W, S) for W in (8, 16) for S in (6, 8, 10, 12, 14, 17, 20, 25, 33, 50)]
for W, S in combos:
    for k, v in (("FBR_WIN_W", W), ("FBR_WIN_S", S)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    plan = E.window_plan(nx, horizon, every, dtype=dt_name)
    if plan is None:
        print(W, S, "no plan"); continue
    ts = []
    for it in range(5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        E.run_window(fac, p0_d, nx, n_steps, w["lbc"], w["rbc"], sidx_d, n_src, table_d, horizon, snapshot_every=every, dtype=dt_name)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"W={W} S={S} plan={plan}  {min(ts[1:]):.3f} ms  {nx * n_steps / min(ts[1:]) / 1e6:.1f} G cu/s")

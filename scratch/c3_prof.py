"""One K3 launch on a C3-shaped ensemble with fewer time steps (profiling target). Scratch tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fiberis_b200 import _engine as E
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
w = bench.make_workload("c3", 1)
M, nx, dt = w["M_per_gpu"], w["nx"], w["dt"]
w["n_steps"] = n_steps
table = bench.source_table(w)[:n_steps]
mesh_d = E.to_device(w["mesh"]); zv_d = E.to_device(w["zv"]); zone_d = E.to_device(w["zone"], np.int32)
sidx_d, n_src = E.index_tensor(w["sidx"]); table_d = E.to_device(table); p0_d = E.to_device(np.zeros(nx))
obs_idx_d, n_obs = E.index_tensor(w["obs_idx"])
audit_d, n_audit = E.index_tensor(np.asarray(w["audit"], dtype=np.int64))
fac = E.factorize(*E.assemble_zonal(mesh_d, zv_d, zone_d, M, nx, dt, "Neumann", "Neumann", sidx_d, n_src), check_singular=False)
obs_d = torch.zeros((n_steps + 1, n_obs), dtype=torch.float64, device=mesh_d.device)
for variant in ("x8", "generic"):
    if variant == "generic":
        os.environ["FBR_K3_GENERIC"] = "1"
    for it in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        E.run(fac, p0_d, M, nx, n_steps, "Neumann", "Neumann", sidx_d, n_src, table_d, snapshot_every=1,
              snap_models=audit_d, n_snap_models=n_audit, obs_idx=obs_idx_d, n_obs=n_obs, obs=obs_d)
        b.record(); torch.cuda.synchronize()
    print(variant, "ms", a.elapsed_time(b))

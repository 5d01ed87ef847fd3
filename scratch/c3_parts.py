"""K3 on the C3 ensemble with parts switched off (what does each ingredient cost in the REAL kernel?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fiberis_b200 import _engine as E
n_steps = 1000
w = bench.make_workload("c3", 1)
M, nx, dt = w["M_per_gpu"], w["nx"], w["dt"]
w["n_steps"] = n_steps
table = bench.source_table(w)[:n_steps]
mesh_d = E.to_device(w["mesh"]); zv_d = E.to_device(w["zv"]); zone_d = E.to_device(w["zone"], np.int32)
sidx_d, n_src = E.index_tensor(w["sidx"]); table_d = E.to_device(table); p0_d = E.to_device(np.zeros(nx))
p1_d = E.to_device(np.random.default_rng(0).uniform(0.5, 1.5, nx))
obs_idx_d, n_obs = E.index_tensor(w["obs_idx"])
audit_d, n_audit = E.index_tensor(np.asarray(w["audit"], dtype=np.int64))
obs_d = torch.zeros((n_steps + 1, n_obs), dtype=torch.float64, device=mesh_d.device)
def fac(lbc, rbc, src):
    s, ns = (sidx_d, n_src) if src else (None, 0)
    return E.factorize(*E.assemble_zonal(mesh_d, zv_d, zone_d, M, nx, dt, lbc, rbc, s, ns), check_singular=False)
def t(name, f, lbc, rbc, src, obs, snap, p0=p0_d):
    ms = []
    for it in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        E.run(f, p0, M, nx, n_steps, lbc, rbc, sidx_d if src else None, n_src if src else 0, table_d if src else None,
              snapshot_every=1 if snap else 0, snap_models=audit_d if snap else None, n_snap_models=n_audit if snap else None,
              obs_idx=obs_idx_d if obs else None, n_obs=n_obs if obs else 0, obs=obs_d if obs else None, want_final=not snap)
        b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    print(f"{name:46s} {min(ms):7.3f} ms  -> C3 {min(ms)*5:6.1f} ms  {min(ms)*1e3/n_steps/ (M/592):.3f} us/step/CTA")
for variant in ("x8", "generic"):
    if variant == "generic": os.environ["FBR_K3_GENERIC"] = "1"
    print("==", variant)
    fN = fac("None", "None", True); f0 = fac("Neumann", "Neumann", False); f1 = fac("Neumann", "Neumann", True)
    t("Neumann BCs only (2 zero overrides), p0 random", f0, "Neumann", "Neumann", False, False, False, p1_d)
    t("+ 2 sources", f1, "Neumann", "Neumann", True, False, False)
    t("+ 2 sources + 16 obs", f1, "Neumann", "Neumann", True, True, False)
    t("+ 2 sources + 16 obs + 4 audit snapshots", f1, "Neumann", "Neumann", True, True, True)

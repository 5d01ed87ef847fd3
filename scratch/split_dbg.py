import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fiberis_b200 import _engine as E
w = bench.make_workload("c3", 1)
M, nx, n_steps = w["M_per_gpu"], w["nx"], w["n_steps"]
table = bench.source_table(w)
mesh_d = E.to_device(w["mesh"]); zv_d = E.to_device(w["zv"]); zone_d = E.to_device(w["zone"], np.int32)
sidx_d, n_src = E.index_tensor(w["sidx"]); table_d = E.to_device(table); p0_d = E.to_device(np.zeros(nx))
obs_idx_d, n_obs = E.index_tensor(w["obs_idx"]); obs_d = E.to_device(np.zeros((n_steps + 1, n_obs)))
pick, n_audit = E.index_tensor(np.asarray(w["audit"], dtype=np.int64))
fac = E.factorize(*E.assemble_zonal(mesh_d, zv_d, zone_d, M, nx, 1.0, "Neumann", "Neumann", sidx_d, n_src), check_singular=False)
main, side = torch.cuda.current_stream(), E._copy_stream(mesh_d.device)
def ev(): return torch.cuda.Event(enable_timing=True)
MODE = sys.argv[1] if len(sys.argv) > 1 else "full"
for it in range(3):
    torch.cuda.synchronize()
    e0, e1, e2, e3, e4 = ev(), ev(), ev(), ev(), ev()
    t0 = time.perf_counter()
    e0.record(main)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        fac_a = tuple(f.index_select(0, pick).contiguous() for f in fac)
        e1.record(side)
        if MODE in ("full", "nod2h"):
            kw = dict(obs_idx=obs_idx_d, n_obs=n_obs, obs=obs_d) if MODE == "full" else {}
            snap_a, _, _ = E.run(fac_a, p0_d, n_audit, nx, n_steps, "Neumann", "Neumann", sidx_d, n_src, table_d, snapshot_every=1, **kw)
        elif it == 0:
            snap_a = torch.zeros((4, n_steps + 1, nx), dtype=torch.float64, device="cuda")
        e2.record(side)
        if MODE in ("full", "d2honly"):
            host = E.to_host(snap_a, sync=False)
        e3.record(side)
    t1 = time.perf_counter()
    if MODE != "noreserve": E.reserve_ctas(n_audit)
    _, _, mis = E.run(fac, p0_d, M, nx, n_steps, "Neumann", "Neumann", sidx_d, n_src, table_d, snapshot_every=0,
                      obs_idx=obs_idx_d, n_obs=n_obs, obs=obs_d)
    e4.record(main)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"host: side enqueue {1e3*(t1-t0):.2f} ms, main enqueue {1e3*(t2-t1):.2f}, total {1e3*(t3-t0):.2f} | gpu: gather {e0.elapsed_time(e1):.2f}, audit run {e1.elapsed_time(e2):.2f}, d2h {e2.elapsed_time(e3):.2f}, main done at {e0.elapsed_time(e4):.2f}, side done at {e0.elapsed_time(e3):.2f}")

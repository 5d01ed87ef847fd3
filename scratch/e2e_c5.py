"""Where does the end-to-end time of a C5 solve go?  (host timeline with synchronisation points)"""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fiberis_b200 import _engine as E
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
w = bench.make_single(name)
p = bench.build_single(w)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        p.solve(dt=w["dt"], t_total=w["T"], snapshot_every=w["every"])
    torch.cuda.synchronize(); print("solve total", round((time.perf_counter() - t0) * 1e3, 1), "ms")
# stage by stage
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {label:40s} {(time.perf_counter() - t0) * 1e3:8.2f} ms"); return r
nx, dt, every, n_steps = w["nx"], w["dt"], w["every"], w["n_steps"]
taxis = T("time axis (python loop)", lambda: [w["dt"] * i for i in range(n_steps + 1)])
table = T("source table (np.interp)", lambda: p._source_table(taxis[:-1]))
mesh_d = T("H2D mesh", lambda: E.to_device(w["mesh"]))
D_d = T("H2D diffusivity", lambda: E.to_device(w["D"]))
p0_d = T("H2D initial", lambda: E.to_device(w["initial"][None, :]))
sidx_d, n_src = E.index_tensor(np.asarray(w["sidx"]))
diag = T("assemble", lambda: E.assemble(mesh_d, D_d, 1, nx, dt, w["lbc"], w["rbc"], sidx_d, n_src))
fac = T("factorize (+flag D2H)", lambda: E.factorize(*diag))
table_d = T("H2D table", lambda: E.to_device(table))
hor = T("decay horizon (+D2H)", lambda: E.decay_horizon(fac))
snap = T("run_window", lambda: E.run_window(fac, p0_d, nx, n_steps, w["lbc"], w["rbc"], sidx_d, n_src, table_d, hor, snapshot_every=every)[0])
host = T("D2H snapshots (pinned pool)", lambda: E.to_host(snap[0]))
host2 = T("D2H snapshots again", lambda: E.to_host(snap[0]))

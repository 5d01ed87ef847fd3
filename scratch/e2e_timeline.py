"""Device timeline of PDS1D_Ensemble.solve on C3 (scratch tool): CUDA events around every _engine.run / as_f64 / to_host
call, on the stream the call is issued on.  python scratch/e2e_timeline.py [float32]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fiberis_b200 import _engine as E

dtype = sys.argv[1] if len(sys.argv) > 1 else "float64"
w = bench.make_workload("c3", 1)
ens = bench.build_ensemble(w)
ens.set_observations(w["obs_idx"], np.zeros((w["n_steps"] + 1, len(w["obs_idx"]))))
marks = []


def wrap(name):
    f = getattr(E, name)

    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = f(*a, **k)
        e1.record()
        marks.append((name + (f"(M={a[2]})" if name == "run" else ""), e0, e1))
        return r
    setattr(E, name, g)


for n in ("run", "as_f64", "to_host", "costly_members_first", "to_device"):
    wrap(n)
for it in range(4):
    marks.clear()
    torch.cuda.synchronize()
    start = torch.cuda.Event(enable_timing=True)
    start.record()
    t0 = time.perf_counter()
    ens.solve(dt=1.0, t_total=float(w["n_steps"]), audit_models=w["audit"], dtype=dtype)
    wall = 1e3 * (time.perf_counter() - t0)
    torch.cuda.synchronize()
print(f"{dtype}: solve() wall {wall:.2f} ms; device timeline (ms after the call started):")
for name, e0, e1 in marks:
    a, b = start.elapsed_time(e0), start.elapsed_time(e1)
    if b - a > 0.02:
        print(f"  {name:28s} {a:7.2f} -> {b:7.2f}  ({b - a:.2f})")

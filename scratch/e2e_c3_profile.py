"""Where does PDS1D_Ensemble.solve spend host time on C3?  cProfile of one end-to-end call (scratch tool)."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
dtype = sys.argv[1] if len(sys.argv) > 1 else "float64"  # float32: the fp32 mode
w = bench.make_workload("c3", 1)
ens = bench.build_ensemble(w)
obs = np.zeros((w["n_steps"] + 1, len(w["obs_idx"])))
ens.set_observations(w["obs_idx"], obs)
for it in range(3):
    ens.solve(dt=1.0, t_total=float(w["n_steps"]), audit_models=w["audit"], dtype=dtype)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
ens.solve(dt=1.0, t_total=float(w["n_steps"]), audit_models=w["audit"], dtype=dtype)
torch.cuda.synchronize()
pr.disable()
print(f"total {1e3 * (time.perf_counter() - t0):.2f} ms")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
# event timeline of the two streams: when does the audited members' chain (kernel, widening, copy home) end vs the sweep?
from fiberis_b200 import _engine
side = _engine._copy_stream(torch.device("cuda", 0))
for it in range(3):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter()
    e0.record()
    ens.solve(dt=1.0, t_total=float(w["n_steps"]), audit_models=w["audit"], dtype=dtype)
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    print(f"solve() wall {1e3 * (t1 - t0):.2f} ms, events {e0.elapsed_time(e1):.2f} ms")


"""Host-side profile of PDS1D_Ensemble.solve on C3 in microseconds (scratch tool; cProfile raw stats, 20 calls)."""
import os, sys, cProfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
dtype = sys.argv[1] if len(sys.argv) > 1 else "float64"
w = bench.make_workload("c3", 1)
ens = bench.build_ensemble(w)
ens.set_observations(w["obs_idx"], np.zeros((w["n_steps"] + 1, len(w["obs_idx"]))))
call = lambda: ens.solve(dt=1.0, t_total=float(w["n_steps"]), audit_models=w["audit"], dtype=dtype)
for _ in range(3):
    call()
pr = cProfile.Profile()
N = 20
pr.enable()
for _ in range(N):
    call()
pr.disable()
rows = []
for e in pr.getstats():
    code = e.code
    name = code if isinstance(code, str) else f"{os.path.basename(code.co_filename)}:{code.co_firstlineno}({code.co_name})"
    rows.append((e.inlinetime / N * 1e6, e.totaltime / N * 1e6, e.callcount / N, name))
rows.sort(reverse=True)
print(f"{'inline us':>10s} {'total us':>10s} {'calls':>6s}  function   (per solve() call)")
for r in rows[:32]:
    print(f"{r[0]:10.1f} {r[1]:10.1f} {r[2]:6.1f}  {r[3][:100]}")

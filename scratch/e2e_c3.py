import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
w = bench.make_workload("c3", 1)
ens = bench.build_ensemble(w)
obs = np.zeros((w["n_steps"] + 1, len(w["obs_idx"])))
ens.set_observations(w["obs_idx"], obs)
def t(label, **kw):
    ts = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ens.solve(dt=1.0, t_total=float(w["n_steps"]), **kw)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"{label:40s} {min(ts[1:])*1e3:7.2f} ms")
t("no audit")
t("4 audited, split", audit_models=w["audit"])
t("4 audited, one launch", audit_models=w["audit"], split_audit=False)

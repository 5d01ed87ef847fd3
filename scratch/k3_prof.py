"""Where a member's time goes in K3 on C3 (library built with -DFBR_K3_PROF): python scratch/k3_prof.py"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fiberis_b200 import _cabi
lib = _cabi.load()
lib.fbr_debug_k3_prof.argtypes = [C.c_void_p, C.c_int]
w = bench.make_workload("c3")
ens = bench.build_ensemble(w)
obs = np.random.default_rng(1).uniform(0, 1, (w["n_steps"] + 1, len(w["obs_idx"])))
ens.set_observations(w["obs_idx"], obs)
for it in range(2):
    lib.fbr_debug_k3_prof(None, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ens.solve(dt=w["dt"], t_total=w["n_steps"] * w["dt"], audit_models=w["audit"])
    torch.cuda.synchronize(); print("solve", time.perf_counter() - t0)
buf = (C.c_longlong * (8 * 8 * 4))()
print("rc", lib.fbr_debug_k3_prof(buf, 0))
a = np.array(buf[:]).reshape(8, 8, 4)
for cta in range(8):
    rows = [r for r in a[cta] if r[2] > 0]
    for r in rows[:8]:
        tot = r.sum()
        print(f"CTA {cta * 71}: prologue {r[0]:8d}  block boundaries {r[1]:9d}  step runs {r[2]:10d}  epilogue {r[3]:7d}   "
              f"per step {r[2] / w['n_steps']:.0f} cycles; boundaries {100 * r[1] / tot:.1f} %, prologue {100 * r[0] / tot:.2f} %")

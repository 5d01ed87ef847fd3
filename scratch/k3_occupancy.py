"""K3 step time against the number of co-resident CTAs per SM (C3's members, 5000 steps): python scratch/k3_occupancy.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
w = bench.make_workload("c3")
obs = np.random.default_rng(1).uniform(0, 1, (w["n_steps"] + 1, len(w["obs_idx"])))
for M in (148, 296, 444, 592, 1184):
    ww = dict(w, zv=w["zv"][:M], M=M)
    ens = bench.build_ensemble(ww)
    ens.set_observations(w["obs_idx"], obs)
    ts = []
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ens.solve(dt=w["dt"], t_total=w["n_steps"] * w["dt"])
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t = min(ts)
    print(f"M={M:5d} ({M / 148:.0f} per SM): {t * 1e3:7.2f} ms  -> {t / w['n_steps'] / max(1, M / 592) * 1e6:.3f} us per step and CTA round "
          f"= {t / w['n_steps'] / max(1, M / 592) * 1965e6:.0f} cycles")

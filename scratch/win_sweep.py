"""Calibration sweep of K3W (window warps x steps per block) on the C2 / C5 meshes. Scratch tool."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fiberis_b200 import _engine as E

def main(name, n_steps, combos):
    w = bench.make_single(name)
    nx, dt, every = w["nx"], w["dt"], w["every"]
    t_axis = np.arange(n_steps, dtype=np.float64) * dt
    table = np.ascontiguousarray(np.stack([np.interp(t_axis, st, sd) for st, sd in w["srcs"]], axis=1))
    mesh_d, D_d, p0_d = E.to_device(w["mesh"]), E.to_device(w["D"]), E.to_device(w["initial"][None, :])
    sidx_d, n_src = E.index_tensor(np.asarray(w["sidx"]))
    table_d = E.to_device(table)
    fac = E.factorize(*E.assemble(mesh_d, D_d, 1, nx, dt, w["lbc"], w["rbc"], sidx_d, n_src), check_singular=False)
    t0 = time.perf_counter(); hor = E.decay_horizon(fac); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(name, "nx", nx, "horizon", hor, f"horizon time {1e3*(t1-t0):.2f} ms (first call)")
    t0 = time.perf_counter(); hor = E.decay_horizon(fac); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"horizon time {1e3*(t1-t0):.2f} ms")
    for W, s in combos:
        os.environ["FBR_WIN_W"] = str(W); os.environ["FBR_WIN_S"] = str(s)
        plan = E.window_plan(nx, hor, every)
        if plan is None:
            print(W, s, "unsupported"); continue
        ts = []
        for it in range(3):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = E.run_window(fac, p0_d, nx, n_steps, w["lbc"], w["rbc"], sidx_d, n_src, table_d, hor, snapshot_every=every)[0]
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
            del out
        ms = min(ts)
        print(f"W={W:2d} s={plan['steps_per_block']:3d} own={plan['owned']:5d} win={plan['windows']:6d}  {ms:9.3f} ms  "
              f"{ms*1e3/n_steps:8.3f} us/step  {nx*n_steps/ms/1e6:8.2f} Gcu/s")

if __name__ == "__main__":
    which = sys.argv[1]
    if which == "c2":
        main("c2", 2000, [(8, s) for s in (8, 10, 11, 12, 13, 14, 17, 20)] + [(16, s) for s in (20, 25, 30, 33)])
    else:
        main("c5", 200, [(W, s) for W in (8, 16) for s in (1, 2, 4, 5, 10, 20, 25)])

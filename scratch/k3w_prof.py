"""Phase timeline of K3W windows (library built with -DFBR_K3W_PROF): python scratch/k3w_prof.py c2|c5"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fiberis_b200 import _cabi
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.make_single(wl)
n = 200
w = dict(w, n_steps=n, T=n * w["dt"])
p = bench.build_single(w)
p.solve(dt=w["dt"], t_total=w["T"], snapshot_every=100, stream_rows=False)
print(p.long_mesh_path)
lib = _cabi.load()
buf = (C.c_longlong * (8 * 8 * 8))()
lib.fbr_debug_window_prof.argtypes = [C.c_void_p]
print("rc", lib.fbr_debug_window_prof(buf))
a = np.array(buf[:]).reshape(8, 8, 8)
names = ["factors in registers", "source list + descriptors", "scan multipliers + barrier", "state loaded (after the grid dependency)",
         "next window staged", "time loop", "written back"]
steps = p.long_mesh_path["steps_per_block"]
for cta in range(8):
    t = a[cta]
    valid = [i for i in range(8) if t[i, 7] > 0 and t[i, 0] > 0]
    if not valid:
        continue
    d = np.diff(t[valid], axis=1)
    per = np.median(d, axis=0)
    print(f"CTA {cta * 17}: windows {len(valid)}  cycles:", {names[k]: int(per[k]) for k in range(7)}, " total", int(np.median(t[valid][:, 7] - t[valid][:, 0])),
          f" per step {per[5] / steps:.0f} (last block may be shorter than {steps} steps)")

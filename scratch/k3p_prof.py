"""Phase timeline of K3P tiles (library built with -DFBR_K3P_PROF): python scratch/k3p_prof.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fiberis_b200 import _cabi
w = bench.make_single("c5")
w = dict(w, n_steps=20, T=20 * w["dt"])
p = bench.build_single(w)
p.solve(dt=w["dt"], t_total=w["T"], snapshot_every=20, long_mesh="exact", _stream_above=0)
print(p.long_mesh_path)
lib = _cabi.load()
buf = (C.c_longlong * (8 * 24 * 12))()
lib.fbr_debug_pass_prof.argtypes = [C.c_void_p]
print("rc", lib.fbr_debug_pass_prof(buf))
a = np.array(buf[:]).reshape(8, 24, 12)
names = ["top", "data arrived", "read out + next copy issued", "forward done", "aggregate published", "carry from the right", "stores issued", "Ef_next + end"]
for cta in range(8):
    t = a[cta]
    valid = [i for i in range(24) if t[i, 7] > 0 and t[i, 0] > 0]
    if len(valid) < 4:
        continue
    tv = t[valid]
    print("   read-out split: rows taken", int(np.median(tv[2:, 8] - tv[2:, 1])), " barrier", int(np.median(tv[2:, 9] - tv[2:, 8])), " copies issued", int(np.median(tv[2:, 2] - tv[2:, 9])))
    d = np.diff(t[valid][:, :8], axis=1)
    per = np.median(d[2:], axis=0)
    gap = np.median(t[valid][1:, 0] - t[valid][:-1, 7])
    total = np.median(t[valid][1:, 0] - t[valid][:-1, 0])
    print(f"CTA {cta * 37}: tiles {len(valid)}  median cycles per phase:", {names[k + 1]: int(per[k]) for k in range(7)}, "loop gap", int(gap), "tile period", int(total))

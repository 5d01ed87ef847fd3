"""Phases of the short-model kernel's time step on C1 (library built with -DFBR_K3G_PROF): python scratch/c1_prof.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fiberis_b200 import _cabi
w = bench.make_single("c1")
p = bench.build_single(w)
for _ in range(2):
    p.solve(dt=w["dt"], t_total=w["T"])
lib = _cabi.load()
buf = (C.c_longlong * 10)()
lib.fbr_debug_k3g_prof.argtypes = [C.c_void_p]
print("rc", lib.fbr_debug_k3g_prof(buf))
names = ["pointer step + prefetch + override", "forward: chunk + 5 stages + carry shuffle", "forward barrier", "forward table + apply",
         "backward: chunk + 5 stages + carry shuffle", "backward barrier", "backward table + apply", "objective", "snapshot row + loop end", "loop top"]
n = w["n_steps"]
tot = sum(buf[:])
for k in range(10):
    print(f"{names[k]:45s} {buf[k] / n:8.1f} cycles per step")
print(f"{'total':45s} {tot / n:8.1f} cycles per step = {tot / n / 1965:.3f} us")

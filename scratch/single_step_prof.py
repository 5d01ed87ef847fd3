"""Where a member's time goes in the one-step use of K3 (65536 x 512, n_steps = 1; -DFBR_K3_PROF build): python scratch/single_step_prof.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fiberis_b200 import _cabi, _engine as E
lib = _cabi.load()
lib.fbr_debug_k3_prof.argtypes = [C.c_void_p, C.c_int]
M, nx = 65536, 512
rng = np.random.default_rng(7)
mesh_d = E.to_device(np.cumsum(rng.uniform(0.5, 2.0, nx)))
zone_d = E.to_device(((np.arange(nx) * 8) // nx).astype(np.int32), np.int32)
zv_d = E.to_device(10.0 ** rng.uniform(-1, 1, (M, 8)))
sidx_d, n_src = E.index_tensor(np.array([nx // 4, 3 * nx // 4]))
fac = E.factorize(*E.assemble_zonal(mesh_d, zv_d, zone_d, M, nx, 1.0, "Neumann", "Neumann", sidx_d, n_src), check_singular=False)
p = torch.rand((M, nx), dtype=torch.float64, device=mesh_d.device)
table_d = E.to_device(np.ones((1, 2)))
for it in range(3):
    lib.fbr_debug_k3_prof(None, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    E.run(fac, p, M, nx, 1, "Neumann", "Neumann", sidx_d, n_src, table_d, snapshot_every=0, want_final=True)
    b.record(); torch.cuda.synchronize()
    print("ms", a.elapsed_time(b))
buf = (C.c_longlong * (8 * 8 * 4))()
lib.fbr_debug_k3_prof(buf, 0)
a = np.array(buf[:]).reshape(8, 8, 4)
for cta in range(8):
    for r in a[cta][:4]:
        if r[2] > 0:
            print(f"CTA {cta * 71}: prologue {r[0]:6d}  block boundary {r[1]:6d}  the step {r[2]:6d}  epilogue {r[3]:6d}   total {r.sum():6d} cycles")

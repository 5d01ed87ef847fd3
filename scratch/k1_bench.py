"""Assembly (K1) and factorisation (K1b) throughput on ensemble shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fiberis_b200 import _engine as E
for M, nx in [(65536, 512), (4096, 1024), (125000, 512)]:
    rng = np.random.default_rng(0)
    mesh_d = E.to_device(np.cumsum(rng.uniform(0.5, 2.0, nx))); zone_d = E.to_device(((np.arange(nx) * 8) // nx).astype(np.int32), np.int32)
    zv_d = E.to_device(10.0 ** rng.uniform(-1, 1, (M, 8))); sidx_d, n_src = E.index_tensor(np.array([nx // 4, 3 * nx // 4]))
    def t(fn):
        ts = []
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        return min(ts), r
    ta, diag = t(lambda: E.assemble_zonal(mesh_d, zv_d, zone_d, M, nx, 1.0, "Neumann", "Neumann", sidx_d, n_src))
    tf, fac = t(lambda: E.factorize(*diag, check_singular=False))
    print(f"M={M:7d} nx={nx:5d}: assemble {ta:7.3f} ms ({M*nx*24/ta/1e6:7.1f} GB/s written)  factorize {tf:7.3f} ms ({M*nx*48/tf/1e6:7.1f} GB/s r+w)")

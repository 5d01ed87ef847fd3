"""One batched implicit step (factors reused, n_steps = 1..8): achieved HBM GB/s at 40 B per cell-update
(24 B factors + 8 B p^n in + 8 B p^{n+1} out) — the step-at-a-time use of K3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fiberis_b200 import _engine as E
for M, nx in [(65536, 512), (32768, 1024), (262144, 256)]:
    rng = np.random.default_rng(0)
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx)); zone = ((np.arange(nx) * 8) // nx).astype(np.int32)
    zv = 10.0 ** rng.uniform(-1, 1, (M, 8))
    mesh_d = E.to_device(mesh); zv_d = E.to_device(zv); zone_d = E.to_device(zone, np.int32)
    sidx_d, n_src = E.index_tensor(np.array([nx // 4, 3 * nx // 4]))
    fac = E.factorize(*E.assemble_zonal(mesh_d, zv_d, zone_d, M, nx, 1.0, "Neumann", "Neumann", sidx_d, n_src), check_singular=False)
    p = torch.rand((M, nx), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for n_steps in (1, 2, 8):
        table_d = E.to_device(np.ones((n_steps, 2)))
        ts = []
        for it in range(4):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = E.run(fac, p, M, nx, n_steps, "Neumann", "Neumann", sidx_d, n_src, table_d, snapshot_every=0, want_final=True)[1]
            b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = min(ts)
        byts = M * nx * (24 + 16)  # per launch: factors + state in/out once
        print(f"M={M:7d} nx={nx:5d} n_steps={n_steps}  {ms:7.3f} ms  HBM {byts/ms/1e6:7.1f} GB/s  {M*nx*n_steps/ms/1e6:8.1f} G cell-updates/s")

"""Cost of recording a snapshot every step: 592 models (one round), all recorded vs none."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fiberis_b200 import _engine as E
n_steps = 500
w = bench.make_workload("c3", 1)
M, nx, dt = 592, w["nx"], w["dt"]
table = bench.source_table(w)[:n_steps]
mesh_d = E.to_device(w["mesh"]); zv_d = E.to_device(w["zv"][:M]); zone_d = E.to_device(w["zone"], np.int32)
sidx_d, n_src = E.index_tensor(w["sidx"]); table_d = E.to_device(table); p0_d = E.to_device(np.zeros(nx))
f = E.factorize(*E.assemble_zonal(mesh_d, zv_d, zone_d, M, nx, dt, "Neumann", "Neumann", sidx_d, n_src), check_singular=False)
for variant in ("x8", "generic"):
    if variant == "generic": os.environ["FBR_K3_GENERIC"] = "1"
    for every in (0, 1, 10):
        ms = []
        for it in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = E.run(f, p0_d, M, nx, n_steps, "Neumann", "Neumann", sidx_d, n_src, table_d, snapshot_every=every, want_final=True)
            b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b)); del out
        print(variant, "snapshot_every", every, f"{min(ms):.3f} ms  {min(ms)*1e3/n_steps:.3f} us/step")

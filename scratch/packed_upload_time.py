"""Host time of the packed upload of C3's small inputs (scratch tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fiberis_b200 import _engine as E
rng = np.random.default_rng(0)
items = [(rng.random(1024), np.float64), (np.array([256, 768]), np.int64), (rng.random((4096, 8)), np.float64),
         (np.arange(1024) // 128, np.int32), (np.zeros(1024), np.float64), (rng.random((5000, 2)), np.float64),
         (np.array([0, 1, 2, 3]), np.int64), (np.arange(16), np.int64), (rng.random((5001, 16)), np.float64), None]
for _ in range(5):
    E.to_device_packed(items)
torch.cuda.synchronize()
for name, f in (("packed", lambda: E.to_device_packed(items)),
                ("one by one", lambda: [None if it is None else E.to_device(it[0], it[1]) for it in items])):
    t0 = time.perf_counter()
    for _ in range(200):
        f()
    host = (time.perf_counter() - t0) / 200
    torch.cuda.synchronize()
    print(f"{name:12s} {1e6 * host:7.1f} us of host time per call")

"""Bare batched tridiagonal solve (K2): achieved GB/s on the 40 B/unknown figure of SURVEY 8(d)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fiberis_b200 import _engine as E, _cabi
def run(M, nx, variant, name):
    g = torch.Generator(device="cuda").manual_seed(1)
    dl = -torch.rand((M, nx), dtype=torch.float64, device="cuda", generator=g)
    du = -torch.rand((M, nx), dtype=torch.float64, device="cuda", generator=g)
    d = 1.0 - dl - du
    b = torch.rand((M, nx), dtype=torch.float64, device="cuda", generator=g)
    ts = []
    for it in range(4):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); x = E.solve_tridiag(dl, d, du, b, variant=variant, check_singular=False); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    # residual check
    r = d * x; r[:, 1:] += dl[:, 1:] * x[:, :-1]; r[:, :-1] += du[:, :-1] * x[:, 1:]
    err = float((r - b).abs().max())
    ms = min(ts)
    print(f"{name:8s} M={M:7d} nx={nx:5d}  {ms:8.3f} ms  {M*nx*40/ms/1e6:8.1f} GB/s (40 B/unknown)  {M*nx/ms/1e6:7.2f} G unknowns/s  resid {err:.1e}")
shapes = [(65536, 512), (4096, 1024), (32768, 1024), (1048576, 64), (524288, 128), (256, 4096), (8192, 4096), (16384, 256), (262144, 256), (16384, 2048), (65536, 500), (65536, 511)]
variants = [(_cabi.SOLVER_THOMAS, "thomas"), (_cabi.SOLVER_PCR, "pcr"), (_cabi.SOLVER_REG, "reg")]
if len(sys.argv) > 1:  # e.g. `k2_bench.py reg 65536x512 32768x1024`
    variants = [v for v in variants if v[1] in sys.argv[1].split(",")]
    if len(sys.argv) > 2: shapes = [tuple(int(q) for q in a.split("x")) for a in sys.argv[2:]]
for M, nx in shapes:
    for v, n in variants:
        try: run(M, nx, v, n)
        except Exception as ex: print(n, M, nx, "ERR", str(ex)[:80])

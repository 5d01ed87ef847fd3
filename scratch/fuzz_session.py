"""One long mixed session (as a user's process would be): fixed-dt single models and ensembles in fp64 / fp32, adaptive
runs, an occasional million-cell mesh (parallel uploads, streamed rows) and bare solves, in random order — shakes out
state carried between calls (pinned pool, upload lanes, reserved CTAs, allocator reuse).  python scratch/fuzz_session.py [n] [seed]"""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pds_oracle as O
from tests import _models as Mo, _golden as G
from tests._fuzz import run_case, run_adaptive_case
from fiberis_b200 import _engine as E

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 500
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
fails, t0, keep = 0, time.time(), []


def big_mesh():
    nx = int(rng.choice([1_050_000, 1_300_001]))
    c = Mo.random_case(int(rng.integers(1 << 30)), nx, n_src=4, T=3.0, knots=5)
    p = Mo.from_fixture(c)
    every = int(rng.choice([1, 2, 6]))
    with contextlib.redirect_stdout(io.StringIO()):
        p.solve(dt=0.5, t_total=3.0, snapshot_every=every)
    want, taxis = O.solve_fixed(c["mesh"], c["diffusivity"], c["initial"], c["lbc"], c["rbc"], c["sourceidx"], c["sources"],
                                0.0, 0.5, 3.0)
    keep.append(p.snapshot)  # results stay alive for a while: the pool must not hand their buffers out again
    return G.rel_max(p.snapshot, want[::every]), f"big mesh nx={nx} every={every}", 1e-10


def bare_solve():
    M, nx = int(rng.choice([1, 7, 300, 5000])), int(rng.choice([3, 50, 64, 333, 1024, 4096, 5000]))
    dl = -rng.uniform(0.1, 1.0, (M, nx)); du = -rng.uniform(0.1, 1.0, (M, nx)); d = 1.0 - dl - du
    b = rng.uniform(-1, 1, (M, nx))
    x = E.solve_tridiag(E.to_device(dl), E.to_device(d), E.to_device(du), E.to_device(b)).cpu().numpy()
    m = int(rng.integers(M))
    return G.rel_max(x[m], O.solve_tridiag(dl[m], d[m], du[m], b[m])), f"bare solve {M}x{nx}", 1e-12


for case in range(n_cases):
    kind = rng.choice(["f64", "f64", "f64", "f32", "adaptive", "bare", "big"], p=[.3, .2, .1, .15, .1, .13, .02])
    try:
        if kind == "f64": err, desc = run_case(rng); gate = 1e-10
        elif kind == "f32": err, desc = run_case(rng, "float32"); gate = 1e-5
        elif kind == "adaptive": err, desc = run_adaptive_case(rng); gate = 1e-9
        elif kind == "bare": err, desc, gate = bare_solve()
        else: err, desc, gate = big_mesh()
        assert err <= gate, f"{desc}: {err:.2e}"
        if len(keep) > 3: keep.pop(0)
    except Exception as ex:
        fails += 1
        print("FAIL", kind, type(ex).__name__, str(ex)[:900].replace("\n", " | "))
print(f"{n_cases} mixed cases, {fails} failures, {time.time() - t0:.0f} s")

"""One small run through every kernel family (target of compute-sanitizer memcheck / racecheck)."""
import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import _models as Mo
from tests.test_gpu_ensemble import make_ensemble

def quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()

c = Mo.random_case(1, 100, n_src=2, T=4.0); p = Mo.from_fixture(c); quiet(lambda: p.solve(dt=0.5, t_total=4.0)); print("K3 generic", p.snapshot.shape)
p = Mo.from_fixture(c); quiet(lambda: p.solve(dt=0.5, t_total=4.0, dtype="float32")); print("K3 fp32", p.snapshot.shape)
c = Mo.random_case(2, 700, n_src=3, T=20.0); p = Mo.from_fixture(c); quiet(lambda: p.solve(dt=0.5, t_total=20.0, snapshot_every=3)); print("K3 x8", p.snapshot.shape)
e, cc = make_ensemble(3, 9, 600, 40, n_obs=5); e.set_observations(cc["obs_idx"], np.zeros((41, 5)), np.ones(5))
quiet(lambda: e.solve(dt=1.0, t_total=40.0, audit_models=[0, 8])); print("K3 x8 ensemble", e.snapshot.shape, e.misfit[:2])
c = Mo.random_case(4, 9000, n_src=4, T=3.0)
p = Mo.from_fixture(c); quiet(lambda: p.solve(dt=0.5, t_total=3.0, long_mesh="windowed", snapshot_every=2)); print("K3W", p.long_mesh_path)
p = Mo.from_fixture(c); quiet(lambda: p.solve(dt=0.5, t_total=3.0, long_mesh="exact")); print("K3L", p.long_mesh_path)
c2 = Mo.random_case(44, 40000, n_src=4, T=2.0)
p = Mo.from_fixture(c2); quiet(lambda: p.solve(dt=0.5, t_total=1.0, long_mesh="exact", _stream_above=0)); print("K3S", p.long_mesh_path)
c = Mo.random_case(5, 200, n_src=2, T=5.0); p = Mo.from_fixture(c); quiet(lambda: p.solve(optimizer=True, dt_init=0.05, t_total=5.0)); print("adaptive device", p.snapshot.shape)
p = Mo.from_fixture(c); quiet(lambda: p.solve(optimizer=True, dt_init=0.05, t_total=2.0, adaptive_on_device=False)); print("adaptive host (K2, K5)", p.snapshot.shape)
c = Mo.random_case(6, 40, n_src=1, T=2.0); p = Mo.from_fixture(c); quiet(lambda: p.solve(dt=0.5, t_total=2.0, mode="explicit")); print("jacobi", p.snapshot.shape)
quiet(lambda: e.solve_adaptive(dt_init=0.05, t_total=3.0, audit_models=[1])); print("ensemble adaptive", e.adaptive["accepted"][:3])
print("done")

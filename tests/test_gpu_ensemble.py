"""Ensemble path on the GPU: per-member pressures against the oracle (= the reference looped over
members), fused misfit against its NumPy definition, and size-independent properties at the full
C3 shape of BASELINE.json (4096 models x 1024 cells)."""
import numpy as np
import numpy.testing as npt
import pytest

from oracle import pds_oracle as O
from tests import _golden as G
from tests import _models as Mo

pytestmark = pytest.mark.gpu


def make_ensemble(seed, M, nx, n_steps, zones=8, n_obs=5, dt=1.0, dense=False):
    from fiberis_b200.simulator.core.ensemble import PDS1D_Ensemble
    rng = np.random.default_rng(seed)
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    zone = (np.arange(nx) * zones) // nx
    zv = 10 ** rng.uniform(-1, 1, (M, zones))
    T = n_steps * dt
    srcs = [(np.linspace(0, T, 16), rng.uniform(1, 10, 16)) for _ in range(2)]
    sidx = [nx // 4, (3 * nx) // 4]
    e = PDS1D_Ensemble()
    e.set_mesh(mesh)
    e.set_source(Mo.sources_of(srcs))
    e.set_sourceidx(sidx)
    e.set_bcs("Neumann", "Neumann")
    e.set_initial(np.zeros(nx))
    if dense:
        e.set_diffusivity(zv[:, zone])
    else:
        e.set_zonal_diffusivity(zv, zone)
    obs_idx = np.linspace(1, nx - 2, n_obs).astype(int)
    return e, dict(mesh=mesh, zone=zone, zv=zv, srcs=srcs, sidx=sidx, obs_idx=obs_idx, T=T, dt=dt)


@pytest.mark.parametrize("M,nx,dense", [(37, 200, False), (9, 1024, False), (20, 96, True), (300, 64, False)])
def test_ensemble_members_and_misfit_vs_oracle(M, nx, dense):
    n_steps = 48
    e, c = make_ensemble(11 + nx, M, nx, n_steps, dense=dense)
    ref0, taxis = O.solve_fixed(c["mesh"], c["zv"][0][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                                c["srcs"], 0.0, c["dt"], c["T"])
    obs = ref0[:, c["obs_idx"]] + 0.01
    w = np.linspace(0.5, 2.0, len(c["obs_idx"]))
    e.set_observations(c["obs_idx"], obs, w)
    audit = [0, M // 2, M - 1]
    mis = e.solve(dt=c["dt"], t_total=c["T"], audit_models=audit)
    npt.assert_array_equal(e.taxis, taxis)
    assert mis.shape == (M,) and e.snapshot.shape == (3, n_steps + 1, nx)
    for k, m in enumerate(audit):
        want, _ = O.solve_fixed(c["mesh"], c["zv"][m][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                                c["srcs"], 0.0, c["dt"], c["T"])
        assert G.rel_max(e.snapshot[k], want) <= 1e-10
    for m in range(0, M, max(1, M // 7)):
        want, _ = O.solve_fixed(c["mesh"], c["zv"][m][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                                c["srcs"], 0.0, c["dt"], c["T"])
        assert mis[m] == pytest.approx(O.ensemble_misfit(want, c["obs_idx"], obs, w), rel=1e-9)


def test_ensemble_member_equals_single_model_api():
    e, c = make_ensemble(3, 5, 300, 30)
    e.solve(dt=1.0, t_total=c["T"], audit_models=[2])
    p = e.member(2)
    p.solve(dt=1.0, t_total=c["T"])
    npt.assert_array_equal(e.snapshot[0], p.snapshot)  # same kernels, same arithmetic: bit-equal


def test_ensemble_fp32_mode():
    e, c = make_ensemble(4, 6, 256, 40)
    e.solve(dt=1.0, t_total=c["T"], audit_models=[1], dtype="float32")
    want, _ = O.solve_fixed(c["mesh"], c["zv"][1][c["zone"]], np.zeros(256), "Neumann", "Neumann", c["sidx"], c["srcs"],
                            0.0, 1.0, c["T"])
    assert G.rel_max(e.snapshot[0], want) <= 1e-5


def test_c3_full_shape_properties():
    """BASELINE.json configs[2] at full width (4096 models x 1024 cells), 200 steps:
    (1) audited members match the oracle, (2) the member that generated the observations has
    misfit exactly 0 (deterministic kernel), (3) linearity: scaling every source by 4 scales every
    pressure by exactly 4 (power-of-two scaling commutes with fp64 rounding), so misfits against
    4x observations scale by exactly 16."""
    M, nx, n_steps = 4096, 1024, 200
    e, c = make_ensemble(303, M, nx, n_steps, n_obs=16)
    e.solve(dt=1.0, t_total=c["T"], audit_models=[0])
    obs = e.snapshot[0][:, c["obs_idx"]].copy()
    e.set_observations(c["obs_idx"], obs)
    audit = [0, 1234, 4095]
    mis = e.solve(dt=1.0, t_total=c["T"], audit_models=audit)
    assert mis[0] == 0.0 and np.all(mis[1:] > 0) and np.all(np.isfinite(mis))
    for k, m in enumerate(audit):
        want, _ = O.solve_fixed(c["mesh"], c["zv"][m][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                                c["srcs"], 0.0, 1.0, c["T"])
        assert G.rel_max(e.snapshot[k], want) <= 1e-10
    snap1 = e.snapshot.copy()
    for s in e.source:
        s.data = s.data * 4.0
    e.set_observations(c["obs_idx"], obs * 4.0)
    mis4 = e.solve(dt=1.0, t_total=c["T"], audit_models=audit)
    npt.assert_array_equal(e.snapshot, snap1 * 4.0)
    npt.assert_array_equal(mis4, mis * 16.0)


def test_sharded_blocks_equal_full_run():
    from fiberis_b200.simulator.core.ensemble import shard_bounds
    e, c = make_ensemble(8, 101, 128, 25)
    ref0, _ = O.solve_fixed(c["mesh"], c["zv"][0][c["zone"]], np.zeros(128), "Neumann", "Neumann", c["sidx"],
                            c["srcs"], 0.0, 1.0, c["T"])
    e.set_observations(c["obs_idx"], ref0[:, c["obs_idx"]])
    full = e.solve(dt=1.0, t_total=c["T"]).copy()
    parts = []
    for r in range(4):
        lo, hi = shard_bounds(101, 4, r)
        parts.append(e.solve(dt=1.0, t_total=c["T"], models=(lo, hi)).copy())
    npt.assert_array_equal(np.concatenate(parts), full)

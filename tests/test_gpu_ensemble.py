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


@pytest.mark.parametrize("M,nx,dense", [(37, 200, False), (9, 1024, False), (20, 96, True), (300, 64, False), (77, 24, False),
                                         (50, 12, False), (33, 130, True), (5, 40, False)])
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
    e.solve(dt=1.0, t_total=c["T"], audit_models=[2], on_chip_assembly=False)
    p = e.member(2)
    p.solve(dt=1.0, t_total=c["T"], on_chip_assembly=False)
    npt.assert_array_equal(e.snapshot[0], p.snapshot)  # same kernels (K1 + K1b + K3), same arithmetic: bit-equal
    # default, both APIs: the matrices are assembled and factorised inside the run kernel (scan-based pivots:
    # last-bit differences against the sequential factorisation)
    e.solve(dt=1.0, t_total=c["T"], audit_models=[2])
    assert G.rel_max(e.snapshot[0], p.snapshot) <= 1e-13
    q = e.member(2)
    q.solve(dt=1.0, t_total=c["T"])
    assert G.rel_max(q.snapshot, p.snapshot) <= 1e-13


@pytest.mark.parametrize("nx,lbc,rbc", [(256, "Neumann", "Neumann"), (300, "Dirichlet", "Neumann"), (777, "Neumann", "Dirichlet"),
                                        (1024, "Dirichlet", "Neumann"), (2049, "Dirichlet", "Dirichlet"), (4096, "Neumann", "Neumann")])
def test_on_chip_assembly_matches_three_kernel_path_and_oracle(nx, lbc, rbc):
    """fbr_run_zonal_f64: K1's assembly + a register factorisation inside the run kernel — every team width, padding
    rows, every boundary type, sources on and next to chunk / warp boundaries; singular members are flagged."""
    M, n_steps = 7, 40
    e, c = make_ensemble(4000 + nx, M, nx, n_steps, n_obs=5, dense=nx in (300, 2049))
    e.set_bcs(lbc, rbc)
    e.set_sourceidx([255 if nx > 300 else 8, 256 if nx > 300 else 9][:len(c["sidx"])] + list(c["sidx"][2:]))
    rng = np.random.default_rng(nx)
    e.set_observations(c["obs_idx"], rng.uniform(-1, 1, (n_steps + 1, 5)))
    mis = e.solve(dt=1.0, t_total=c["T"], audit_models=[0, 3, 6]).copy()
    snap = e.snapshot.copy()
    mis3 = e.solve(dt=1.0, t_total=c["T"], audit_models=[0, 3, 6], on_chip_assembly=False)
    assert G.rel_max(snap, e.snapshot) <= 1e-12
    npt.assert_allclose(mis, mis3, rtol=1e-11)
    for k, m in enumerate([0, 3, 6]):
        want, _ = O.solve_fixed(c["mesh"], c["zv"][m][c["zone"]], np.zeros(nx), lbc, rbc, list(e.sourceidx), c["srcs"], 0.0, 1.0,
                                c["T"])
        assert G.rel_max(snap[k], want) <= 1e-10
    if nx == 300:  # a 'None' boundary leaves a zero row (matbuilder.py:49-69): scipy raises in the reference
        e.set_bcs("None", "Neumann")
        with pytest.raises(np.linalg.LinAlgError):
            e.solve(dt=1.0, t_total=c["T"])


def test_ensemble_fp32_mode():
    e, c = make_ensemble(4, 6, 256, 40)
    e.solve(dt=1.0, t_total=c["T"], audit_models=[1], dtype="float32")
    want, _ = O.solve_fixed(c["mesh"], c["zv"][1][c["zone"]], np.zeros(256), "Neumann", "Neumann", c["sidx"], c["srcs"],
                            0.0, 1.0, c["T"])
    assert G.rel_max(e.snapshot[0], want) <= 1e-5


@pytest.mark.parametrize("nx", [48, 100, 200])
def test_ensemble_fp32_mode_short_meshes(nx):
    """fp32 mode on the member-group kernel (<= 128 cells) and on the generic kernel (200 cells)."""
    e, c = make_ensemble(40 + nx, 33, nx, 40)
    e.solve(dt=1.0, t_total=c["T"], audit_models=[0, 32], dtype="float32")
    for k, m in enumerate([0, 32]):
        want, _ = O.solve_fixed(c["mesh"], c["zv"][m][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"], c["srcs"],
                                0.0, 1.0, c["T"])
        assert G.rel_max(e.snapshot[k], want) <= 1e-5


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


def test_members_handed_out_on_demand_equal_the_fixed_deal(monkeypatch):
    """Sweeps that record nothing hand the members beyond the first deal to whichever CTA is free (a ticket word per
    launch; `FBR_K3_STATIC_DEAL` keeps the fixed deal k, k + G, ..), the costly ones first (`member_order`): which CTA
    integrates a member, and when, changes no bit of its misfit, in either precision, and repeated launches (fresh
    ticket words) agree with each other."""
    M, nx, n_steps = 3000, 512, 80  # 8 CTAs of 2 warps per SM: 1184 CTAs < M
    e, c = make_ensemble(77, M, nx, n_steps, n_obs=6)
    ref0, _ = O.solve_fixed(c["mesh"], c["zv"][5][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                            c["srcs"], 0.0, 1.0, c["T"])
    e.set_observations(c["obs_idx"], ref0[:, c["obs_idx"]])
    for dtype in ("float64", "float32"):
        monkeypatch.setenv("FBR_K3_STATIC_DEAL", "1")
        fixed = e.solve(dt=1.0, t_total=c["T"], dtype=dtype).copy()
        monkeypatch.delenv("FBR_K3_STATIC_DEAL")
        for _ in range(3):  # default: costly members first (fbr_run_opts.member_order), handed out on demand
            npt.assert_array_equal(e.solve(dt=1.0, t_total=c["T"], dtype=dtype), fixed)
        monkeypatch.setenv("FBR_NO_MEMBER_ORDER", "1")  # on demand, in index order
        npt.assert_array_equal(e.solve(dt=1.0, t_total=c["T"], dtype=dtype), fixed)
        monkeypatch.delenv("FBR_NO_MEMBER_ORDER")
        assert np.all(np.isfinite(fixed)) and (dtype == "float32" or fixed[5] <= 1e-16)


def test_rank_members_orders_by_decreasing_cost_class():
    """fbr_rank_members_f64 (the order a sweep hands its members out in): a permutation whose half-octave cost classes
    never increase; zero / negative / NaN figures come last."""
    import torch
    from fiberis_b200 import _cabi, _engine
    rng = np.random.default_rng(5)
    for M in (1, 7, 1000, 4096, 70001):
        cost = 10 ** rng.uniform(-6, 6, M)
        if M > 100:
            cost[[3, 50, 77]] = [0.0, -2.0, np.nan]
        cost_d = _engine.to_device(cost)
        order_d = torch.empty(M, dtype=torch.int32, device=cost_d.device)
        _cabi.check(_cabi.load().fbr_rank_members_f64(_cabi.ptr(cost_d), M, _cabi.ptr(order_d), _cabi.current_stream()))
        order = order_d.cpu().numpy()
        npt.assert_array_equal(np.sort(order), np.arange(M))
        c = cost[order]
        good = c > 0
        assert not np.any(good[np.argmin(good):]) or good.all()  # the unusable figures are a suffix
        cls = np.frombuffer(c[good].tobytes(), dtype=np.int64) >> 51
        assert np.all(np.diff(cls) <= 0)


@pytest.mark.parametrize("layout", ["time_major", "cell_major"])
def test_fp32_audited_rows_leave_the_kernel_as_doubles(layout):
    """fp32 sweeps: the audited members' launch stores its rows widened (fbr_run_opts.snap_f64) instead of a widening
    pass behind it; the doubles are exactly the floats the single-launch path records."""
    M, nx, n_steps = 600, 1024, 70
    e, c = make_ensemble(91, M, nx, n_steps, n_obs=4)
    audit = [599, 3, 3, 200]
    e.solve(dt=1.0, t_total=c["T"], audit_models=audit, dtype="float32", split_audit=False, result_layout=layout)
    want = e.snapshot.copy()
    e.solve(dt=1.0, t_total=c["T"], audit_models=audit, dtype="float32", result_layout=layout)
    assert e.snapshot.dtype == np.float64 and e.snapshot.shape == (4, n_steps + 1, nx)
    npt.assert_array_equal(e.snapshot, want)
    npt.assert_array_equal(e.snapshot, e.snapshot.astype(np.float32).astype(np.float64))
    ref, _ = O.solve_fixed(c["mesh"], c["zv"][200][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                           c["srcs"], 0.0, 1.0, c["T"])
    assert G.rel_max(e.snapshot[3], ref) <= 1e-5


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


# ---- the fp64 x8 run kernel: staged blocks of per-step scalars, observation trace, model schedule ----
@pytest.mark.parametrize("n_steps", [1, 31, 32, 33, 64, 65, 100])
@pytest.mark.parametrize("every", [1, 3, 40])
def test_x8_block_boundaries_and_snapshot_rows(n_steps, every):
    """Source values and observations are staged 32 steps at a time and residuals are summed at block
    boundaries: every alignment of run length, block length and snapshot interval must agree with the
    oracle, for snapshots and for the fused misfit."""
    M, nx = 5, 1024
    e, c = make_ensemble(900 + n_steps, M, nx, n_steps, n_obs=7)
    rng = np.random.default_rng(n_steps)
    obs = rng.uniform(-1, 1, (n_steps + 1, 7))
    w = rng.uniform(0.5, 2.0, 7)
    e.set_observations(c["obs_idx"], obs, w)
    mis = e.solve(dt=c["dt"], t_total=c["T"], audit_models=[1, 4], snapshot_every=every)
    for k, m in enumerate([1, 4]):
        want, _ = O.solve_fixed(c["mesh"], c["zv"][m][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                                c["srcs"], 0.0, c["dt"], c["T"])
        rows = want[::every][: 1 + n_steps // every]
        assert e.snapshot[k].shape == rows.shape  # a run shorter than the interval keeps the initial row only
        if n_steps >= every:
            assert G.rel_max(e.snapshot[k], rows) <= 1e-10
        else:
            npt.assert_array_equal(e.snapshot[k], rows)
        assert mis[m] == pytest.approx(O.ensemble_misfit(want, c["obs_idx"], obs, w), rel=1e-9)


def test_x8_recorded_models_schedule_with_duplicates_and_many_models():
    """Recorded models are integrated first, by the CTAs at the end of the grid, and skipped in the
    normal deal; duplicates in the list fill both slots; every other model is still integrated
    exactly once (its misfit appears) whatever M is relative to the grid."""
    nx, n_steps = 512, 40
    for M, audit in [(700, [699, 0, 7, 7, 123]), (40, [39, 38, 0]), (3, [2, 2, 2, 1]), (1300, list(range(0, 1300, 100)))]:
        e, c = make_ensemble(50 + M, M, nx, n_steps, n_obs=4)
        obs = np.zeros((n_steps + 1, 4))
        e.set_observations(c["obs_idx"], obs)
        mis = e.solve(dt=c["dt"], t_total=c["T"], audit_models=audit).copy()
        snap = e.snapshot.copy()
        plain = e.solve(dt=c["dt"], t_total=c["T"])  # no recorded models: the plain deal
        npt.assert_array_equal(mis, plain)
        assert np.all(mis > 0)
        for k, m in enumerate(audit):
            want, _ = O.solve_fixed(c["mesh"], c["zv"][m][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                                    c["srcs"], 0.0, c["dt"], c["T"])
            assert G.rel_max(snap[k], want) <= 1e-10
            assert mis[m] == pytest.approx(O.ensemble_misfit(want, c["obs_idx"], obs, None), rel=1e-9)


def test_x8_several_overrides_and_observations_in_one_thread():
    """Two sources / two observed cells inside one thread's 8-cell chunk (and a source next to a
    boundary row) take the generic per-thread path of the x8 kernel."""
    from fiberis_b200.simulator.core.ensemble import PDS1D_Ensemble
    rng = np.random.default_rng(77)
    M, nx, n_steps = 6, 600, 70
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    zone = (np.arange(nx) * 4) // nx
    zv = 10 ** rng.uniform(-1, 1, (M, 4))
    T = float(n_steps)
    sidx = [1, 2, 300, 301, 303, 598]
    srcs = [(np.linspace(0, T, 9), rng.uniform(-5, 5, 9)) for _ in sidx]
    e = PDS1D_Ensemble()
    e.set_mesh(mesh); e.set_source(Mo.sources_of(srcs)); e.set_sourceidx(sidx); e.set_bcs("Dirichlet", "Neumann")
    e.set_initial(rng.uniform(0, 1, nx)); e.set_zonal_diffusivity(zv, zone)
    obs_idx = np.array([8, 9, 15, 100, 302, 304, 599])
    obs = rng.uniform(-1, 1, (n_steps + 1, len(obs_idx)))
    w = rng.uniform(0.5, 2.0, len(obs_idx))
    e.set_observations(obs_idx, obs, w)
    mis = e.solve(dt=1.0, t_total=T, audit_models=[0, 5])
    for k, m in enumerate([0, 5]):
        want, _ = O.solve_fixed(mesh, zv[m][zone], e.initial, "Dirichlet", "Neumann", sidx, srcs, 0.0, 1.0, T)
        assert G.rel_max(e.snapshot[k], want) <= 1e-10
        assert mis[m] == pytest.approx(O.ensemble_misfit(want, obs_idx, obs, w), rel=1e-9)


# ---- adaptive time stepping per ensemble member -------------------------------------
def test_ensemble_adaptive_every_member_has_its_own_time_axis():
    M, nx = 24, 300
    e, c = make_ensemble(61, M, nx, 30)
    out = e.solve_adaptive(dt_init=0.01, t_total=30.0, tol=2e-3, max_dt=2.0, audit_models=[0, 7, 23])
    assert out["final_state"].shape == (M, nx) and np.all(out["final_time"] >= 30.0)
    assert len(set(out["accepted"].tolist())) > 1  # members with different diffusivity take different paths
    for k, m in enumerate([0, 7, 23]):
        snap, taxis, n_it = O.solve_adaptive(c["mesh"], c["zv"][m][c["zone"]], np.zeros(nx), "Neumann", "Neumann",
                                             c["sidx"], c["srcs"], 0.0, 0.01, 30.0, tol=2e-3, max_dt=2.0)
        assert out["taxis"][k].shape == taxis.shape and out["iterations"][m] == n_it
        npt.assert_allclose(out["taxis"][k], taxis, rtol=1e-9, atol=0)
        assert G.rel_max(out["snapshot"][k], snap) <= 1e-9
        npt.assert_array_equal(out["final_state"][m], out["snapshot"][k][-1])
        assert out["accepted"][m] == len(taxis) - 1
    # a member equals the single-model API run of the same member
    p = e.member(7)
    p.solve(optimizer=True, dt_init=0.01, t_total=30.0, tol=2e-3, max_dt=2.0)
    npt.assert_array_equal(out["snapshot"][1], p.snapshot)
    npt.assert_array_equal(out["taxis"][1], p.taxis)


# ---- sweep driver ---------------------------------------------------------------
def test_sweep_finds_the_generating_candidate_and_refines():
    from fiberis_b200.simulator.optimizer import sweep
    M, nx, n_steps = 64, 256, 60
    e, c = make_ensemble(71, M, nx, n_steps, zones=3, n_obs=6)
    truth = np.array([0.3, -0.4, 0.1])
    want, _ = O.solve_fixed(c["mesh"], (10.0 ** truth)[c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                            c["srcs"], 0.0, c["dt"], c["T"])
    theta = sweep.random_candidates(M, [-1, -1, -1], [1, 1, 1], seed=5)
    theta[17] = truth
    e.set_zonal_diffusivity(10.0 ** theta, c["zone"])  # observations = what candidate 17 produces on the device
    e.solve(dt=c["dt"], t_total=c["T"], audit_models=[17])
    assert G.rel_max(e.snapshot[0], want) <= 1e-10
    e.set_observations(c["obs_idx"], e.snapshot[0][:, c["obs_idx"]].copy())
    out = sweep.run_sweep(e, theta, c["zone"], c["dt"], c["T"])
    assert out["best"] == 17 and out["best_misfit"] == 0.0
    npt.assert_array_equal(out["best_theta"], truth)
    assert np.all(np.diff(out["misfit"][out["order"]]) >= 0)
    ref = sweep.refine_sweep(e, c["zone"], [-1, -1, -1], [1, 1, 1], c["dt"], c["T"], n_models=256, rounds=5, seed=9)
    best = [h[0] for h in ref["history"]]
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:])) and best[-1] < 0.05 * best[0]
    # the zones that hold the sources are well determined by the data (the one between them is not)
    assert np.max(np.abs(ref["best_theta"] - truth)[[0, 2]]) < 0.05


def test_tiled_factorisation_is_bit_identical_to_the_plain_recurrence(monkeypatch):
    """K1b for many models streams through transposing shared-memory tiles; per model it performs the
    same operations in the same order as the thread-per-model kernel."""
    import torch
    from fiberis_b200 import _engine as E
    for M, nx in [(64, 33), (100, 512), (257, 1000), (1000, 31)]:
        rng = np.random.default_rng(M + nx)
        mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
        zone = (np.arange(nx) * 4) // nx
        zv = 10.0 ** rng.uniform(-1, 1, (M, 4))
        sidx_d, n_src = E.index_tensor(np.array([nx // 3]))
        diag = E.assemble_zonal(E.to_device(mesh), E.to_device(zv), E.to_device(zone, np.int32), M, nx, 0.7, "Neumann",
                                "Dirichlet", sidx_d, n_src)
        tiled = torch.stack(E.factorize(*diag, parallel=False)).cpu().numpy()
        monkeypatch.setenv("FBR_FACTORIZE_PLAIN", "1")
        plain = torch.stack(E.factorize(*diag, parallel=False)).cpu().numpy()
        monkeypatch.delenv("FBR_FACTORIZE_PLAIN")
        npt.assert_array_equal(tiled, plain)
    # zero pivots are still flagged per model
    dl = np.zeros((70, 40)); d = np.ones((70, 40)); du = np.zeros((70, 40)); d[5, 17] = 0.0
    with pytest.raises(np.linalg.LinAlgError):
        E.factorize(E.to_device(dl), E.to_device(d), E.to_device(du))


def test_c4_full_member_count_properties():
    """BASELINE.json configs[3] at its full member count (1,000,000 x 512 cells), 24 steps: sampled members' misfits
    equal the banded C oracle's, the generating member's misfit is exactly 0, and the blocks of an 8-way shard are
    bit-equal to the corresponding part of the unsharded sweep."""
    from fiberis_b200.simulator.core.ensemble import shard_bounds
    from oracle import c_oracle as CO
    M, nx, n_steps = 1_000_000, 512, 24
    e, c = make_ensemble(404, M, nx, n_steps, n_obs=16)
    e.solve(dt=1.0, t_total=c["T"], audit_models=[0], models=(0, 8))
    obs = e.snapshot[0][:, c["obs_idx"]].copy()
    e.set_observations(c["obs_idx"], obs)
    mis = e.solve(dt=1.0, t_total=c["T"], audit_models=[0, 999_999]).copy()
    assert mis.shape == (M,) and mis[0] == 0.0 and np.all(np.isfinite(mis)) and np.count_nonzero(mis) >= M - 1
    table = np.stack([np.interp(np.arange(n_steps) * 1.0, t, d) for t, d in c["srcs"]], axis=1)  # sampled at step start
    pick = np.array([1, 2, 77_777, 500_000, 999_998, 999_999])
    want, _ = CO.run_ensemble(c["mesh"], c["zv"][pick], c["zone"], np.zeros(nx), n_steps, 1.0, "Neumann", "Neumann",
                              np.asarray(c["sidx"]), table, c["obs_idx"], obs)
    npt.assert_allclose(mis[pick], want, rtol=1e-9)
    for r in (0, 3, 7):
        lo, hi = shard_bounds(M, 8, r)
        part = e.solve(dt=1.0, t_total=c["T"], models=(lo, hi))
        npt.assert_array_equal(part, mis[lo:hi])


@pytest.mark.parametrize("M,nx,every", [(70, 40, 3), (9, 64, 1), (130, 100, 7), (41, 16, 2), (300, 128, 5)])
def test_member_group_kernel_records_every_member(M, nx, every):
    """Short meshes: groups of lanes per member (run_group_kernel) straight through the engine — snapshots of ALL
    members (snap_models = None), a snapshot interval, the final state, a source on a boundary row and a duplicated
    source index, member counts that leave the last warp partly filled."""
    from fiberis_b200 import _engine as E
    rng = np.random.default_rng(M + nx)
    n_steps = 23
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    D = 10 ** rng.uniform(-1, 1, (M, nx))
    p0 = rng.uniform(0, 1, (M, nx))
    sidx = np.array([nx // 3, nx - 1, nx // 3])  # later duplicate wins; the last one overrides the Dirichlet row
    table = rng.uniform(-5, 5, (n_steps, 3))
    sidx_d, n_src = E.index_tensor(sidx)
    fac = E.factorize(*E.assemble(E.to_device(mesh), E.to_device(D), M, nx, 0.5, "Neumann", "Dirichlet", sidx_d, n_src))
    snap, final, _ = E.run(fac, E.to_device(p0), M, nx, n_steps, "Neumann", "Dirichlet", sidx_d, n_src, E.to_device(table),
                           snapshot_every=every, want_final=True)
    snap, final = snap.cpu().numpy(), final.cpu().numpy()
    assert snap.shape == (M, n_steps // every + 1, nx)
    for m in range(0, M, max(1, M // 6)):
        dl, d, du = O.assemble_tridiag(mesh, D[m], 0.5, "Neumann", "Dirichlet", list(sidx))
        p = p0[m].copy()
        rows = [p.copy()]
        for n in range(n_steps):
            b = p.copy(); b[0] = 0.0; b[-1] = 0.0
            for k, s_ in enumerate(sidx):
                b[s_] = table[n, k]
            p = O.solve_tridiag(dl, d, du, b)
            if (n + 1) % every == 0:
                rows.append(p.copy())
        assert G.rel_max(snap[m], np.stack(rows)) <= 1e-10, m
        assert G.rel_max(final[m], p) <= 1e-10


@pytest.mark.parametrize("nx,dt", [(1024, 30.0), (200, 30.0), (64, 30.0), (1024, 1e-4)])
def test_stiff_and_nearly_explicit_ensembles(nx, dt):
    """alpha = D dt / dx^2 up to ~1.2e3 (multipliers close to one: slow decay, every scan stage and cross-warp carry
    matters) and down to ~1e-6 (multipliers ~1e-6: products underflow towards zero) on the on-chip-assembly path, the
    partly filled x8 warp and the member-group kernel."""
    M, n_steps = 12, 30
    e, c = make_ensemble(77 + nx, M, nx, n_steps, dt=dt)
    e.set_initial(np.random.default_rng(nx).uniform(0, 1, nx))
    e.solve(dt=dt, t_total=c["T"], audit_models=[0, 5, 11])
    for k, m in enumerate([0, 5, 11]):
        want, taxis = O.solve_fixed(c["mesh"], c["zv"][m][c["zone"]], e.initial, "Neumann", "Neumann", c["sidx"], c["srcs"],
                                    0.0, dt, c["T"])
        assert e.snapshot[k].shape == want.shape
        assert G.rel_max(e.snapshot[k], want) <= 1e-10


# ---- observations on their own time axis: the reference's history-matching misfit -------------------------
def _sampled_ensemble(g, M, rng, dtype_dense=True):
    """An ensemble whose member 0 is the fixture's model; the others have perturbed diffusivities."""
    from fiberis_b200.simulator.core.ensemble import PDS1D_Ensemble
    D = np.tile(g["diffusivity"], (M, 1))
    D[1:] *= 10 ** rng.uniform(-0.3, 0.3, (M - 1, 1))
    e = PDS1D_Ensemble()
    e.set_mesh(g["mesh"])
    e.set_source(Mo.sources_of(g["sources"]))
    e.set_sourceidx(g["sourceidx"])
    e.set_bcs(g["lbc"], g["rbc"])
    e.set_initial(g["initial"])
    e.set_diffusivity(D)
    return e, D


def test_fused_misfit_matches_the_reference_misfit_calculator():
    """Member 0 is the model the reference's UNMODIFIED misfit_calculator was run on (tests/golden/
    misfit_reference_320x96.npz: irregular observed taxis, observed start_time 3.5 s after the simulation's, baseline
    sample 1, window [1:-1], channel weights): the misfit K3 fuses into its time loop must reproduce that value."""
    g = G.load("misfit_reference_320x96")
    rng = np.random.default_rng(5)
    M = 9
    e, D = _sampled_ensemble(g, M, rng)
    e.set_observations(g["obs_cells"], g["obs"], weights=g["obs_weights"], taxis=g["obs_taxis"] + float(g["obs_time_offset"]),
                       baseline_index=1, window=(1, -1))
    mis = e.solve(dt=g["dt"], t_total=g["t_total"], audit_models=[0])
    assert mis[0] == pytest.approx(float(g["misfit_total"]), rel=1e-10)
    assert G.rel_max(e.snapshot[0][:, g["obs_cells"]], g["sim_channels"]) <= 1e-10
    for m in range(1, M):
        snap, taxis = O.solve_fixed(g["mesh"], D[m], g["initial"], g["lbc"], g["rbc"], g["sourceidx"], g["sources"], g["t0"],
                                    g["dt"], g["t_total"])
        want = O.sampled_misfit(snap, taxis, g["obs_cells"], g["obs"], g["obs_taxis"], g["obs_weights"],
                                time_offset=float(g["obs_time_offset"]))
        assert mis[m] == pytest.approx(want, rel=1e-10)


def test_set_observed_data_aligns_start_times_like_data1d_interpolate():
    import datetime

    from fiberis_b200.analyzer.Data2D.core2D import Data2D
    g = G.load("misfit_reference_320x96")
    e, D = _sampled_ensemble(g, 3, np.random.default_rng(6))
    observed = Data2D(data=np.ascontiguousarray(g["obs"].T), taxis=g["obs_taxis"], daxis=g["mesh"][g["obs_cells"]],
                      start_time=Mo.START + datetime.timedelta(seconds=float(g["obs_time_offset"])))
    e.set_observed_data(observed, g["obs_cells"], weights=g["obs_weights"])  # reference defaults: baseline 1, window [1:-1]
    mis = e.solve(dt=g["dt"], t_total=g["t_total"])
    assert mis[0] == pytest.approx(float(g["misfit_total"]), rel=1e-10)


@pytest.mark.parametrize("nx,n_steps,n_obs,J,dtype", [
    (1024, 150, 16, 97, "float64"),    # C3 width, several staged blocks, fewer samples than steps
    (520, 70, 40, 400, "float64"),     # observations much denser than the steps (more samples per block than staged)
    (200, 33, 5, 20, "float64"),       # one warp, a partly filled one
    (24, 40, 3, 30, "float64"),        # short mesh: the call must pick the 64-byte-per-thread kernel by itself
    (2500, 20, 7, 15, "float64"),      # 10 warps
    (1024, 130, 9, 60, "float32"),     # fp32 mode
])
def test_sampled_misfit_fuzz(nx, n_steps, n_obs, J, dtype):
    rng = np.random.default_rng(nx + J)
    M = 6
    e, c = make_ensemble(3 * nx + J, M, nx, n_steps, dense=(nx == 520))
    cells = np.sort(rng.choice(np.arange(nx), size=n_obs, replace=False))
    if n_obs >= 5:
        cells[1] = cells[0] + 1 if cells[0] + 1 not in cells[2:] else cells[1]  # two observed cells in one thread
        cells = np.unique(cells)
        n_obs = len(cells)
    T = c["T"]
    obs_taxis = np.sort(rng.uniform(0.0, T, J))
    obs_taxis[0] = 0.0                      # exactly the initial state
    obs_taxis[-1] = T                       # exactly the last state
    obs_taxis[J // 2] = c["dt"] * (n_steps // 2)  # exactly on a step time
    obs_taxis = np.sort(obs_taxis)
    obs = rng.normal(0, 1, (J, n_obs))
    w = rng.uniform(0.5, 2.0, n_obs)
    for base, window in ((1, (1, -1)), (None, None), (0, (2, J - 3))):
        e.set_observations(cells, obs, weights=w, taxis=obs_taxis, baseline_index=base, window=window)
        mis = e.solve(dt=c["dt"], t_total=T, dtype=dtype)
        for m in range(M):
            snap, taxis = O.solve_fixed(c["mesh"], c["zv"][m][c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                                        c["srcs"], 0.0, c["dt"], T)
            want = O.sampled_misfit(snap, taxis, cells, obs, obs_taxis, w, baseline_index=base, window=window)
            assert mis[m] == pytest.approx(want, rel=1e-9 if dtype == "float64" else 2e-3), (m, base, window)


def test_sampled_misfit_outside_the_simulated_range_is_nan_like_the_reference():
    e, c = make_ensemble(77, 4, 300, 20)
    cells = np.array([10, 150, 280])
    obs_taxis = np.array([0.0, 2.5, 7.0, 19.0, 25.0, 30.0])  # the last two lie behind the end of the run (T = 20)
    obs = np.zeros((6, 3))
    e.set_observations(cells, obs, taxis=obs_taxis, baseline_index=1, window=(1, -1))  # sample 4 (outside) is inside the window
    assert np.all(np.isnan(e.solve(dt=c["dt"], t_total=c["T"])))
    e.set_observations(cells, obs, taxis=obs_taxis, baseline_index=1, window=(1, 4))   # now it is not
    mis = e.solve(dt=c["dt"], t_total=c["T"])
    assert np.all(np.isfinite(mis))
    snap, taxis = O.solve_fixed(c["mesh"], c["zv"][2][c["zone"]], np.zeros(300), "Neumann", "Neumann", c["sidx"], c["srcs"], 0.0,
                                c["dt"], c["T"])
    assert mis[2] == pytest.approx(O.sampled_misfit(snap, taxis, cells, obs, obs_taxis, window=(1, 4)), rel=1e-9)
    e.set_observations(cells, obs, taxis=obs_taxis, baseline_index=3, window=(1, 4))  # a baseline behind the window's start
    with pytest.raises(ValueError):
        e.solve(dt=c["dt"], t_total=c["T"])


# ---- per-member scan dispatch ------------------------------------------------------------------------------------------
def test_scan_modes_are_exercised_and_agree_with_the_exact_scans():
    """C3-shaped members take the truncated scans (neighbour-warp carry; 4 stages when every 16-chunk product is below
    2^-80); a smooth high-diffusivity mesh keeps the exact ones.  Truncated and exact results agree to rounding."""
    import torch

    from fiberis_b200 import _engine as E
    rng = np.random.default_rng(3)
    M, nx, n_steps = 64, 1024, 40
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    zone = ((np.arange(nx) * 8) // nx).astype(np.int32)
    zv = 10 ** rng.uniform(-1, 1, (M, 8))
    zv[0] = 3000.0   # decay horizon far beyond a warp: the exact scans must be kept for this member
    table = rng.uniform(1, 10, (n_steps, 2))
    mesh_d, zv_d, zone_d = E.to_device(mesh), E.to_device(zv), E.to_device(zone, np.int32)
    sidx_d, n_src = E.index_tensor(np.array([256, 768]))
    table_d, p0_d = E.to_device(table), E.to_device(np.zeros(nx))
    out = {}
    for name, eps in (("default", 0.0), ("exact", 1.0)):
        modes = torch.full((M,), -1, dtype=torch.int32, device=mesh_d.device)
        snap, _, _ = E.run(None, p0_d, M, nx, n_steps, "Neumann", "Neumann", sidx_d, n_src, table_d, snapshot_every=10,
                           zonal=(mesh_d, zv_d, zone_d, 1.0, None), scan_log2_eps=eps, scan_mode_out=modes)
        out[name] = (snap.cpu().numpy(), modes.cpu().numpy())
    m_def, m_ex = out["default"][1], out["exact"][1]
    assert np.all(m_ex == 0)
    assert m_def[0] == 0 and set(m_def[1:]) <= {1, 2} and (m_def == 2).sum() > 0 and (m_def == 1).sum() > 0
    assert G.rel_max(out["default"][0], out["exact"][0]) <= 1e-13
    want, _ = O.solve_fixed(mesh, zv[5][zone], np.zeros(nx), "Neumann", "Neumann", [256, 768],
                            [(np.arange(n_steps + 1.0), np.append(table[:, k], table[-1, k])) for k in range(2)], 0.0, 1.0, float(n_steps))
    assert G.rel_max(out["default"][0][5], want[::10]) <= 1e-10


# ---- cell-major snapshots: the (cell, time) layout of Data2D written by the kernel ---------------------------
@pytest.mark.parametrize("nx,M", [(1024, 5), (200, 3), (64, 40), (24, 9)])
def test_cell_major_snapshots_are_the_transpose(nx, M):
    from fiberis_b200 import _engine as E
    rng = np.random.default_rng(nx)
    n_steps = 37
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    D = 10 ** rng.uniform(-1, 1, (M, nx))
    sidx_d, n_src = E.index_tensor(np.array([nx // 3]))
    table_d = E.to_device(rng.uniform(1, 10, (n_steps, 1)))
    p0_d = E.to_device(rng.uniform(0, 1, (M, nx)))
    fac = E.factorize(*E.assemble(E.to_device(mesh), E.to_device(D), M, nx, 0.5, "Dirichlet", "Neumann", sidx_d, n_src))
    args = (fac, p0_d, M, nx, n_steps, "Dirichlet", "Neumann", sidx_d, n_src, table_d)
    rows, _, _ = E.run(*args, snapshot_every=3)
    cols, _, _ = E.run(*args, snapshot_every=3, cell_major=True)
    assert cols.shape == (M, nx, rows.shape[1])
    npt.assert_array_equal(cols.cpu().numpy(), rows.cpu().numpy().transpose(0, 2, 1))


def test_ensemble_cell_major_audit_snapshots():
    e, c = make_ensemble(21, 600, 512, 40)
    e.solve(dt=c["dt"], t_total=c["T"], audit_models=[3, 599, 3])
    ref = e.snapshot.copy()
    e.solve(dt=c["dt"], t_total=c["T"], audit_models=[3, 599, 3], result_layout="cell_major")
    assert e.snapshot.shape == ref.shape
    npt.assert_array_equal(e.snapshot, ref)
    assert np.transpose(e.snapshot, (0, 2, 1)).flags["C_CONTIGUOUS"]


def test_sweep_with_observations_on_their_own_time_axis():
    """The sweep driver on the reference's misfit: gauge channels recorded on an irregular time axis (offset from the
    simulation's), baseline sample 1, window [1:-1].  The generating candidate has the smallest misfit — not 0: the
    observed samples are interpolated values of ITS states, which the kernel reproduces to rounding."""
    from fiberis_b200.simulator.optimizer import sweep
    M, nx, n_steps = 48, 512, 80
    e, c = make_ensemble(91, M, nx, n_steps, zones=3, n_obs=4)
    truth = np.array([0.2, -0.3, 0.4])
    want, taxis = O.solve_fixed(c["mesh"], (10.0 ** truth)[c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                                c["srcs"], 0.0, c["dt"], c["T"])
    rng = np.random.default_rng(2)
    obs_taxis = np.sort(rng.uniform(0.5, c["T"] - 0.5, 33))
    obs = np.stack([np.interp(obs_taxis, taxis, want[:, i]) for i in c["obs_idx"]], axis=1)
    obs = obs - obs[1]
    e.set_observations(c["obs_idx"], obs, taxis=obs_taxis, baseline_index=1, window=(1, -1))
    theta = sweep.random_candidates(M, [-1, -1, -1], [1, 1, 1], seed=4)
    theta[5] = truth
    out = sweep.run_sweep(e, theta, c["zone"], c["dt"], c["T"])
    assert out["best"] == 5 and out["best_misfit"] < 1e-20
    for m in (0, 11, 47):
        snap, _ = O.solve_fixed(c["mesh"], (10.0 ** theta[m])[c["zone"]], np.zeros(nx), "Neumann", "Neumann", c["sidx"],
                                c["srcs"], 0.0, c["dt"], c["T"])
        assert out["misfit"][m] == pytest.approx(O.sampled_misfit(snap, taxis, c["obs_idx"], obs, obs_taxis), rel=1e-9)


@pytest.mark.parametrize("M,nx", [(5, 257), (40, 512), (33, 1000), (3, 4096), (17, 2047)])
@pytest.mark.parametrize("lbc,rbc", [("Neumann", "Dirichlet"), ("Dirichlet", "Neumann")])
def test_one_step_with_reused_factors_takes_the_register_resident_step_kernel(M, nx, lbc, rbc, monkeypatch):
    """fbr_run_f64 with n_steps = 1 and nothing recorded but the new state (the step-at-a-time use) is served by K2's
    kernel in its FACTORS mode (csrc/fbr_solve_reg.cu): same answer as the persistent kernel and as the oracle's step,
    with sources on a boundary row, next to each other and listed twice."""
    from fiberis_b200 import _engine as E
    from oracle import pds_oracle as O
    rng = np.random.default_rng(M * 7919 + nx)
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    D = 10 ** rng.uniform(-1, 1, (M, nx))
    p = rng.uniform(-1, 1, (M, nx))
    sidx = np.array([0, 5, 6, nx // 2, nx // 2, nx - 1])
    vals = rng.uniform(-10, 10, (1, len(sidx)))
    mesh_d, D_d, p_d = E.to_device(mesh), E.to_device(D), E.to_device(p)
    sidx_d, n_src = E.index_tensor(sidx)
    fac = E.factorize(*E.assemble(mesh_d, D_d, M, nx, 0.5, lbc, rbc, sidx_d, n_src))
    table_d = E.to_device(vals)
    _, fin, _ = E.run(fac, p_d, M, nx, 1, lbc, rbc, sidx_d, n_src, table_d, snapshot_every=0, want_final=True)
    got = fin.cpu().numpy()
    monkeypatch.setenv("FBR_K3_NO_STEP_KERNEL", "1")
    _, fin2, _ = E.run(fac, p_d, M, nx, 1, lbc, rbc, sidx_d, n_src, table_d, snapshot_every=0, want_final=True)
    assert np.max(np.abs(got - fin2.cpu().numpy())) <= 1e-13 * np.max(np.abs(got))
    for m in (0, M - 1):
        b = O.assemble_rhs(p[m], lbc, rbc, sidx, vals[0])
        want = O.solve_tridiag(*O.assemble_tridiag(mesh, D[m], 0.5, lbc, rbc, sidx), b)
        assert np.max(np.abs(got[m] - want)) <= 1e-12 * np.max(np.abs(want))

"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

TEST INFRASTRUCTURE.  Usage:  python oracle/gen_golden.py   (needs /root/reference mounted).
The fixtures are committed; this script is committed with them so they can be regenerated.
Every fixture stores its full inputs, so the parity tests need neither the reference nor this
script at run time.  Versions used are stored in each file (numpy/scipy).
"""
from __future__ import annotations

import contextlib
import datetime
import io
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
START = datetime.datetime(2020, 1, 1)
VERS = dict(numpy_version=np.__version__, scipy_version=scipy.__version__)


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _mk(R, mesh, D, initial, lbc, rbc, sidx, srcs, t0, multi):
    p = R.pds.PDS1D_MultiSource() if multi else R.pds.PDS1D_SingleSource()
    with _quiet():
        p.set_mesh(np.asarray(mesh, dtype=float))
        ds = [R.Data1D(data=np.asarray(d, float), taxis=np.asarray(t, float), start_time=START) for t, d in srcs]
        p.set_source(ds if multi else ds[0])
        p.set_bcs(lbc, rbc)
        p.set_initial(np.asarray(initial, dtype=float))
        p.set_diffusivity(D if np.isscalar(D) else np.asarray(D, dtype=float))
        p.set_sourceidx(list(sidx) if multi else int(sidx[0]))
        p.set_t0(t0)
    return p


def _save(name, mesh, D, initial, lbc, rbc, sidx, srcs, t0, multi, snapshot, taxis, rows=None, **extra):
    nx = len(mesh)
    Dv = np.full(nx, float(D)) if np.isscalar(D) else np.asarray(D, float)
    rows = np.arange(len(taxis)) if rows is None else np.asarray(rows)
    kn = max(len(t) for t, _ in srcs)
    st = np.full((len(srcs), kn), np.nan)
    sd = np.full((len(srcs), kn), np.nan)
    for k, (t, d) in enumerate(srcs):
        st[k, :len(t)] = t
        sd[k, :len(d)] = d
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), mesh=np.asarray(mesh, float), diffusivity=Dv,
        initial=np.asarray(initial, float), lbc=lbc, rbc=rbc, sourceidx=np.asarray(sidx, np.int64),
        src_taxis=st, src_data=sd, t0=float(t0), multi=bool(multi), taxis=np.asarray(taxis, float),
        rows=rows, snapshot_rows=np.asarray(snapshot)[rows], snapshot_sum=np.asarray(snapshot).sum(axis=1),
        **extra, **VERS)
    print(f"{name}: nx={nx} nt={len(taxis) - 1} rows={len(rows)}")


def gauge(rng, T, knots=64, lo=-10, hi=10):
    return np.linspace(0, T, knots), rng.uniform(lo, hi, knots)


def misfit_fixture(R):
    """The history-matching misfit of the reference: ``misfit_calculator``
    (src/fiberis/moose/templates/baseline_model_generator.py:418-520) run UNMODIFIED on a reference solve and a
    synthetic observed Data2D with its own, irregular time axis and a later start_time."""
    from fiberis.moose.templates import baseline_model_generator as B  # type: ignore
    rng = np.random.default_rng(61)
    nx = 320
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    D = 10 ** rng.uniform(-1, 1, nx)
    init = rng.uniform(0, 1, nx)
    T = 48.0
    srcs = [gauge(rng, T, 12, 1.0, 10.0) for _ in range(2)]
    sidx = [80, 240]
    p = _mk(R, mesh, D, init, "Neumann", "Neumann", sidx, srcs, 0.0, True)
    with _quiet():
        p.solve(dt=0.5, t_total=T)
    conv = 0.3048
    sim = R.Data2D(data=np.array(p.snapshot).T.copy(), taxis=np.array(p.taxis, float), daxis=mesh * conv, start_time=START,
                   name="simulated")
    # observed: 5 channels around cell 121 (3 of them weighted), 41 irregular samples, recorded from 3.5 s after the
    # simulation's start_time; values = a perturbed copy of another model's response
    cells = np.arange(119, 124)
    obs_start = START + datetime.timedelta(seconds=3.5)
    obs_taxis = np.sort(rng.uniform(0.0, T - 4.0, 41))
    obs_taxis[0] = 0.0
    snap = np.array(p.snapshot)
    obs = np.stack([np.interp(obs_taxis + 3.5, p.taxis, snap[:, c]) for c in cells]) * 0.9 + rng.normal(0, 0.05, (5, 41))
    obs = obs - obs[:, 1:2]
    observed = R.Data2D(data=obs, taxis=obs_taxis, daxis=mesh[cells].copy(), start_time=obs_start, name="observed")
    weights = np.array([0.5, 2.0, 1.0])
    cwd = os.getcwd()
    import tempfile
    with tempfile.TemporaryDirectory() as tmp, _quiet():
        os.chdir(tmp)
        try:
            total, per_channel = B.misfit_calculator(weight_matrix=weights, sim_fracture_center_ind=121,
                                                     observed_data_fracture_center_ind=2, simulated_data=sim,
                                                     observed_data=observed, save_path=tmp, return_misfit_per_channel=True)
        finally:
            os.chdir(cwd)
    used_cells = np.array([120, 121, 122])  # channels start at observed index 2 - 3 // 2 = 1
    _save("misfit_reference_320x96", mesh, D, init, "Neumann", "Neumann", sidx, srcs, 0.0, True, p.snapshot, p.taxis,
          rows=[0, 1, 48, 96], dt=0.5, t_total=T, obs_cells=used_cells, obs=obs[1:4].T.copy(), obs_taxis=obs_taxis,
          obs_time_offset=3.5, obs_weights=weights, baseline_index=1, window=np.array([1, -1]),
          sim_channels=snap[:, used_cells].copy(), misfit_total=float(total), misfit_per_channel=np.asarray(per_channel))
    print("misfit_reference_320x96: total", total, per_channel)


def main():
    R = _refshim.import_reference()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "misfit":
        misfit_fixture(R)
        return

    # ---- C1 (BASELINE.json configs[0]; SURVEY 8(d) seed 101): fully through the reference
    mesh = np.linspace(0, 199, 200)
    srcs = [(np.array([0., 500., 1000.]), np.array([0., 10., 10.]))]
    p = _mk(R, mesh, 1.0, np.zeros(200), "Neumann", "Neumann", [100], srcs, 0.0, False)
    with _quiet():
        p.solve(dt=1.0, t_total=1000.0)
    _save("c1_single_200x1000", mesh, 1.0, np.zeros(200), "Neumann", "Neumann", [100], srcs, 0.0, False,
          p.snapshot, p.taxis, rows=[0, 1, 2, 3, 10, 100, 499, 500, 501, 999, 1000], dt=1.0, t_total=1000.0)

    # ---- non-uniform mesh, variable D, Neumann/Dirichlet, 3 sources, random initial
    rng = np.random.default_rng(202)
    nx = 96
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    D = 10 ** rng.uniform(-1, 1, nx)
    init = rng.uniform(0, 1, nx)
    T = 40.0
    srcs = [gauge(rng, T, 16) for _ in range(3)]
    sidx = [9, 48, 77]
    p = _mk(R, mesh, D, init, "Neumann", "Dirichlet", sidx, srcs, 0.0, True)
    with _quiet():
        p.solve(dt=0.5, t_total=T)
    _save("multi_nonuniform_96x80", mesh, D, init, "Neumann", "Dirichlet", sidx, srcs, 0.0, True,
          p.snapshot, p.taxis, dt=0.5, t_total=T)

    # ---- t0 != 0, t_total defaulted from the source, non-dyadic dt (fp64 taxis accumulation)
    rng = np.random.default_rng(7)
    nx = 33
    mesh = np.cumsum(rng.uniform(0.2, 1.0, nx))
    D = 10 ** rng.uniform(-0.5, 0.5, nx)
    srcs = [(np.array([0., 0.3, 0.7, 1.0]), np.array([1., 4., -2., 3.]))]
    p = _mk(R, mesh, D, np.zeros(nx), "Dirichlet", "Neumann", [16], srcs, 2.5, False)
    with _quiet():
        p.solve(dt=0.1)
    _save("single_t0_default_ttotal_33", mesh, D, np.zeros(nx), "Dirichlet", "Neumann", [16], srcs, 2.5, False,
          p.snapshot, p.taxis, dt=0.1, t_total=np.nan)

    # ---- sources sitting ON the boundary rows, 'None' BC on the left made regular by a source
    rng = np.random.default_rng(11)
    nx = 40
    mesh = np.linspace(0, 39, nx)
    srcs = [gauge(rng, 10.0, 8) for _ in range(2)]
    sidx = [0, nx - 1]
    p = _mk(R, mesh, 2.0, rng.uniform(-1, 1, nx), "None", "Neumann", sidx, srcs, 0.0, True)
    with _quiet():
        p.solve(dt=0.25, t_total=10.0)
    _save("multi_boundary_sources_none_bc_40", mesh, 2.0, p.initial, "None", "Neumann", sidx, srcs, 0.0, True,
          p.snapshot, p.taxis, dt=0.25, t_total=10.0)

    # ---- duplicate source index (later wins), adjacent sources
    rng = np.random.default_rng(12)
    nx = 24
    mesh = np.linspace(0, 23, nx)
    srcs = [gauge(rng, 5.0, 6) for _ in range(4)]
    sidx = [5, 6, 5, 20]
    p = _mk(R, mesh, 1.5, np.zeros(nx), "Dirichlet", "Dirichlet", sidx, srcs, 0.0, True)
    with _quiet():
        p.solve(dt=0.5, t_total=5.0)
    _save("multi_duplicate_sources_24", mesh, 1.5, np.zeros(nx), "Dirichlet", "Dirichlet", sidx, srcs, 0.0, True,
          p.snapshot, p.taxis, dt=0.5, t_total=5.0)

    # ---- minimum mesh nx=2 and nx=3
    for nx in (2, 3):
        mesh = np.linspace(0, nx - 1.0, nx)
        srcs = [(np.array([0., 1., 2.]), np.array([1., 2., 0.]))]
        p = _mk(R, mesh, 1.0, np.ones(nx), "Neumann", "Dirichlet", [0], srcs, 0.0, False)
        with _quiet():
            p.solve(dt=0.5, t_total=2.0)
        _save(f"single_tiny_{nx}", mesh, 1.0, np.ones(nx), "Neumann", "Dirichlet", [0], srcs, 0.0, False,
              p.snapshot, p.taxis, dt=0.5, t_total=2.0)

    # ---- a C3-style ensemble member: 1024 cells, zonal D, 2 sources, Neumann/Neumann, 60 steps
    rng = np.random.default_rng(303)
    nx = 1024
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    theta = rng.uniform(-1, 1, 8)
    D = 10 ** theta[(np.arange(nx) * 8) // nx]
    T = 60.0
    srcs = [gauge(rng, T, 64) for _ in range(2)]
    sidx = [256, 768]
    p = _mk(R, mesh, D, np.zeros(nx), "Neumann", "Neumann", sidx, srcs, 0.0, True)
    with _quiet():
        p.solve(dt=1.0, t_total=T)
    _save("c3_member_1024x60", mesh, D, np.zeros(nx), "Neumann", "Neumann", sidx, srcs, 0.0, True,
          p.snapshot, p.taxis, rows=[0, 1, 2, 30, 59, 60], dt=1.0, t_total=T, theta=theta)

    # ---- a larger alpha (stiff) case: alpha up to ~1e3
    rng = np.random.default_rng(13)
    nx = 300
    mesh = np.cumsum(rng.uniform(0.1, 0.3, nx))
    D = 10 ** rng.uniform(0, 1, nx)
    srcs = [gauge(rng, 20.0, 10)]
    p = _mk(R, mesh, D, rng.uniform(0, 1, nx), "Neumann", "Neumann", [150], srcs, 0.0, False)
    with _quiet():
        p.solve(dt=1.0, t_total=20.0)
    _save("single_stiff_300x20", mesh, D, p.initial, "Neumann", "Neumann", [150], srcs, 0.0, False,
          p.snapshot, p.taxis, dt=1.0, t_total=20.0)

    # ---- explicit (Jacobi) mode
    mesh = np.linspace(0, 8, 9)
    srcs = [(np.array([0., 5., 10.]), np.array([0., 10., 10.]))]
    p = _mk(R, mesh, 1.0, np.zeros(9), "Dirichlet", "Dirichlet", [4], srcs, 0.0, False)
    with _quiet():
        p.solve(dt=0.2, t_total=1.0, mode="explicit")
    _save("single_explicit_9", mesh, 1.0, np.zeros(9), "Dirichlet", "Dirichlet", [4], srcs, 0.0, False,
          p.snapshot, p.taxis, dt=0.2, t_total=1.0)

    # ---- adaptive dt (optimizer=True), single and multi, incl. user tol + custom kwargs
    rng = np.random.default_rng(21)
    nx = 50
    mesh = np.cumsum(rng.uniform(0.5, 1.5, nx))
    D = 10 ** rng.uniform(-0.5, 0.5, nx)
    srcs = [(np.array([0., 2., 10., 30.]), np.array([5., 8., 3., 6.]))]
    p = _mk(R, mesh, D, np.full(nx, 1.0), "Neumann", "Neumann", [25], srcs, 0.0, False)
    with _quiet():
        p.solve(optimizer=True, dt_init=0.05, t_total=30.0)
    _save("adaptive_single_50", mesh, D, np.full(nx, 1.0), "Neumann", "Neumann", [25], srcs, 0.0, False,
          p.snapshot, p.taxis, dt_init=0.05, t_total=30.0)
    srcs = [(np.array([0., 2., 10., 30.]), np.array([5., 8., 3., 6.])),
            (np.array([0., 15., 30.]), np.array([-2., 4., 1.]))]
    p = _mk(R, mesh, D, np.full(nx, 0.5), "Dirichlet", "Neumann", [10, 40], srcs, 1.0, True)
    with _quiet():
        p.solve(optimizer=True, dt_init=0.1, t_total=25.0, tol=5e-3, safety_factor=0.8, max_dt=2.0, min_dt=1e-3, p=2)
    _save("adaptive_multi_50_kwargs", mesh, D, np.full(nx, 0.5), "Dirichlet", "Neumann", [10, 40], srcs, 1.0, True,
          p.snapshot, p.taxis, dt_init=0.1, t_total=25.0, tol=5e-3, safety_factor=0.8, max_dt=2.0, min_dt=1e-3, p=2)

    # ---- matbuilder with PML > 0 (dense A, b) single (sigma_max default 2) and multi (default 10)
    rng = np.random.default_rng(31)
    nx = 21
    mesh = np.cumsum(rng.uniform(0.5, 1.5, nx))
    D = 10 ** rng.uniform(-0.5, 0.5, nx)
    srcs = [(np.array([0., 5., 10.]), np.array([0., 10., 10.]))]
    prev = rng.uniform(0, 1, nx)
    p = _mk(R, mesh, D, prev, "Neumann", "Dirichlet", [10], srcs, 0.0, False)
    p.taxis = [2.5]
    p.snapshot = [prev]
    A1, b1 = R.matbuilder.matrix_builder_1d_single_source(p, 0.3, pml_thickness=4.0)
    A2, b2 = R.matbuilder.matrix_builder_1d_single_source(p, 0.3, pml_thickness=3.0, sigma_max=5.0)
    pm = _mk(R, mesh, D, prev, "Dirichlet", "Neumann", [3, 17], srcs * 2, 0.0, True)
    pm.taxis = [2.5]
    pm.snapshot = [prev]
    A3, b3 = R.matbuilder.matrix_builder_1d_multi_source(pm, 0.3, pml_thickness=4.0)
    np.savez_compressed(os.path.join(OUT, "matbuilder_pml_21.npz"), mesh=mesh, diffusivity=D, prev=prev,
                        src_taxis=srcs[0][0], src_data=srcs[0][1], A1=A1, b1=b1, A2=A2, b2=b2, A3=A3, b3=b3,
                        sigma1=R.matbuilder.build_pml_sigma(mesh, 4.0, 2), **VERS)
    print("matbuilder_pml_21")

    # ---- np.interp semantics through Data1D.get_value_by_time on a realistic gauge
    rng = np.random.default_rng(41)
    gt = np.cumsum(rng.uniform(0.5, 3.0, 96))
    gd = rng.integers(1000, 9000, 96).astype(float)
    g = R.Data1D(data=gd, taxis=gt, start_time=START)
    q = np.concatenate([[-5.0, gt[0], gt[-1], gt[-1] + 10], rng.uniform(gt[0], gt[-1], 200), gt[5:10]])
    np.savez_compressed(os.path.join(OUT, "data1d_interp_96.npz"), taxis=gt, data=gd, query=q,
                        value=np.array([g.get_value_by_time(float(t)) for t in q]), **VERS)
    print("data1d_interp_96")

    misfit_fixture(R)


if __name__ == "__main__":
    main()

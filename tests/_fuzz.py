"""One randomly drawn fixed-dt configuration (single model of any length, or an ensemble: zonal / dense, on-chip
assembly on or off, any member-group width) solved on the GPU and compared with the NumPy oracle.  Shared by
tests/test_gpu_fuzz.py and scratch/fuzz_parity.py."""
import contextlib
import io

import numpy as np

from oracle import pds_oracle as O
from tests import _golden as G
from tests import _models as Mo


def run_case(rng, dtype="float64"):
    """Returns (relative max-norm error, description); raises AssertionError on a taxis / misfit mismatch."""
    from fiberis_b200.simulator.core.ensemble import PDS1D_Ensemble
    kind = rng.choice(["single", "single", "ensemble", "ensemble", "long"])  # (round 2: the fp32 mode covers long meshes too)
    if kind == "single":
        nx = int(rng.choice([2, 3, 5, 9, 31, 64, 200, 255, 256, 257, 700, 1024, 2049, 4096]))
    elif kind == "long":
        nx = int(rng.choice([4097, 6000, 20001, 70000, 150003]))
    else:
        nx = int(rng.choice([9, 16, 17, 40, 64, 65, 128, 129, 200, 256, 300, 512, 1000, 2048]))
    lbc, rbc = (str(rng.choice(["Dirichlet", "Neumann"])) for _ in range(2))
    n_src = int(rng.integers(1, 5)) if nx > 6 else 1
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    init = rng.uniform(0, 1, nx) if rng.random() < 0.7 else np.zeros(nx)
    dt = float(rng.choice([0.1, 0.5, 1.0, 1 / 3, 2.5]))
    n_steps = int(rng.integers(1, 70))
    T = dt * n_steps if rng.random() < 0.7 else dt * (n_steps - 0.5)
    t0 = float(rng.choice([0.0, 0.0, 1.5, -2.0]))
    every = int(rng.choice([1, 1, 2, 7]))
    cand = np.arange(0, nx) if nx <= 6 else np.arange(1, nx - 1)
    sidx = rng.choice(cand, size=min(n_src, len(cand)), replace=bool(rng.random() < 0.2)).tolist()
    if rng.random() < 0.15:
        sidx[0] = int(rng.choice([0, nx - 1]))
    srcs = [(np.linspace(0, max(T, dt), 7), rng.uniform(-10, 10, 7)) for _ in sidx]
    desc = f"{kind} nx={nx} {lbc}/{rbc} src={sidx} dt={dt} steps~{n_steps} t0={t0} every={every}"
    if kind in ("single", "long"):
        D = 10 ** rng.uniform(-1, 1, nx)
        p = Mo.model(mesh, D, init, lbc, rbc, sidx, srcs, t0, True)
        with contextlib.redirect_stdout(io.StringIO()):
            p.solve(dt=dt, t_total=T + t0, snapshot_every=every, dtype=dtype)
        want, taxis = O.solve_fixed(mesh, D, init, lbc, rbc, sidx, srcs, t0, dt, T + t0)
        kept = np.arange(0, len(taxis), every)[: 1 + (len(taxis) - 1) // every]
        assert np.array_equal(p.taxis, taxis[kept]), "taxis"
        err = G.rel_max(p.snapshot, want[kept])
    else:
        M = int(rng.choice([1, 3, 8, 33, 70, 300]))
        dense = rng.random() < 0.4
        zones = int(rng.choice([1, 3, 8]))
        zone = (np.arange(nx) * zones) // nx
        zv = 10 ** rng.uniform(-1, 1, (M, zones))
        e = PDS1D_Ensemble()
        e.set_mesh(mesh); e.set_source(Mo.sources_of(srcs)); e.set_sourceidx(sidx); e.set_bcs(lbc, rbc)
        e.set_initial(init); e.t0 = t0
        if dense: e.set_diffusivity(zv[:, zone])
        else: e.set_zonal_diffusivity(zv, zone)
        n_obs = int(rng.integers(1, 6))
        oidx = rng.choice(nx, size=min(n_obs, nx), replace=False)
        taxis_o = O.solve_fixed(mesh, zv[0][zone], init, lbc, rbc, sidx, srcs, t0, dt, T + t0)[1]
        obs = rng.uniform(-1, 1, (len(taxis_o), len(oidx)))
        w = rng.uniform(0.5, 2, len(oidx)) if rng.random() < 0.5 else None
        e.set_observations(oidx, obs, w)
        audit = sorted(set(int(a) for a in rng.choice(M, size=min(3, M), replace=False)))
        mis = e.solve(dt=dt, t_total=T + t0, audit_models=audit, snapshot_every=every, dtype=dtype,
                      on_chip_assembly=bool(rng.random() < 0.7))
        err = 0.0
        for k, m in enumerate(audit):
            want, taxis = O.solve_fixed(mesh, zv[m][zone], init, lbc, rbc, sidx, srcs, t0, dt, T + t0)
            kept = np.arange(0, len(taxis), every)[: 1 + (len(taxis) - 1) // every]
            assert np.array_equal(e.snapshot_taxis, taxis[kept]), "snapshot_taxis"
            err = max(err, G.rel_max(e.snapshot[k], want[kept]))
            ref_mis = O.ensemble_misfit(want, oidx, obs, w)
            mis_tol = 1e-9 if dtype == "float64" else 1e-3
            assert abs(mis[m] - ref_mis) <= mis_tol * max(1.0, abs(ref_mis)), f"misfit {mis[m]} vs {ref_mis}"
    return err, desc


def run_adaptive_case(rng):
    """optimizer=True (step doubling, on the device) for one random model against the oracle's loop: identical
    accept / reject history (same number of rows), times to 1e-9, rows to 1e-9.  Returns (error, description)."""
    nx = int(rng.choice([5, 9, 40, 200, 256, 257, 600, 1024, 1500, 3000]))
    lbc, rbc = (str(rng.choice(["Dirichlet", "Neumann"])) for _ in range(2))
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    D = 10 ** rng.uniform(-1, 1, nx)
    init = rng.uniform(0, 1, nx)
    T = float(rng.choice([5.0, 20.0, 60.0]))
    n_src = int(rng.integers(1, 4)) if nx > 6 else 1
    sidx = sorted(rng.choice(np.arange(1, nx - 1), size=n_src, replace=False).tolist())
    srcs = [(np.linspace(0, T, 9), rng.uniform(-10, 10, 9)) for _ in sidx]
    kw = dict(tol=float(rng.choice([1e-3, 5e-3, 2e-2])), max_dt=float(rng.choice([0.5, 2.0, 40.0])))
    dt_init = float(rng.choice([0.01, 0.05, 0.3]))
    desc = f"adaptive nx={nx} {lbc}/{rbc} src={sidx} T={T} dt_init={dt_init} {kw}"
    p = Mo.model(mesh, D, init, lbc, rbc, sidx, srcs, 0.0, True)
    with contextlib.redirect_stdout(io.StringIO()):
        p.solve(optimizer=True, dt_init=dt_init, t_total=T, **kw)
    snap, taxis, _ = O.solve_adaptive(mesh, D, init, lbc, rbc, sidx, srcs, 0.0, dt_init, T, **kw)
    assert p.taxis.shape == taxis.shape, f"{desc}: {p.taxis.shape} vs {taxis.shape} accepted steps"
    t_err = float(np.max(np.abs(p.taxis - taxis) / np.maximum(np.abs(taxis), 1e-300)))
    s_err = G.rel_max(p.snapshot, snap)
    if t_err > 1e-9 or s_err > 1e-9:
        # The controller amplifies rounding: dt' = dt * 0.9 * sqrt(tol / err) with err a norm of a DIFFERENCE of two nearly
        # equal solutions.  How far do the reference algorithm's OWN accepted times move when the initial state changes by one
        # rounding error?  (tiny models with a loose tol: up to 6e-9, measured on the oracle alone.)  Parity is only meaningful
        # down to that conditioning; anything beyond 20x of it is a real difference.
        init2 = init * (1.0 + 2.2e-16 * rng.choice([-1.0, 1.0], nx))
        snap2, taxis2, _ = O.solve_adaptive(mesh, D, init2, lbc, rbc, sidx, srcs, 0.0, dt_init, T, **kw)
        assert taxis2.shape == taxis.shape, f"{desc}: the oracle's own history changes under a one-ulp perturbation"
        cond_t = float(np.max(np.abs(taxis2 - taxis) / np.maximum(np.abs(taxis), 1e-300)))
        cond_s = G.rel_max(snap2, snap)
        assert t_err <= max(1e-9, 20 * cond_t) and s_err <= max(1e-9, 20 * cond_s), \
            f"{desc}: times {t_err:.2e} (conditioning {cond_t:.2e}), rows {s_err:.2e} (conditioning {cond_s:.2e})"
        return min(s_err, 1e-9), desc + f" [ill-conditioned controller: times {t_err:.1e}, one-ulp sensitivity {cond_t:.1e}]"
    return s_err, desc


def run_misc_case(rng):
    """The less travelled modes of ``solve``: PML damping (pml_thickness / sigma_max), the Jacobi "explicit" mode,
    the host-driven adaptive loop.  Returns (error, gate, description)."""
    kind = str(rng.choice(["pml", "pml", "jacobi", "adaptive_host"]))
    nx = int(rng.choice([5, 12, 64, 200, 300] if kind != "pml" else [12, 64, 200, 257, 1024, 3000, 5000]))
    lbc, rbc = (str(rng.choice(["Dirichlet", "Neumann"])) for _ in range(2))
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    D = 10 ** rng.uniform(-1, 1, nx)
    init = rng.uniform(0, 1, nx)
    n_src = int(rng.integers(1, 3))
    sidx = sorted(rng.choice(np.arange(1, nx - 1), size=n_src, replace=False).tolist())
    T = float(rng.choice([4.0, 10.0]))
    srcs = [(np.linspace(0, T, 6), rng.uniform(-10, 10, 6)) for _ in sidx]
    p = Mo.model(mesh, D, init, lbc, rbc, sidx, srcs, 0.0, True)
    desc = f"{kind} nx={nx} {lbc}/{rbc} src={sidx} T={T}"
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "pml":
            L, smax = float(rng.uniform(1.0, 0.3 * (mesh[-1] - mesh[0]))), float(rng.choice([0.5, 3.0, 10.0]))
            # (dt * sigma_max = 1 on a Neumann row makes its diagonal -1 + dt sigma exactly zero: the un-pivoted elimination
            #  flags it and solve() falls back to the pivoted variant — round 2, tests/test_gpu_parity.py)
            p.solve(dt=0.5, t_total=T, pml_thickness=L, sigma_max=smax)
            want, taxis = O.solve_fixed(mesh, D, init, lbc, rbc, sidx, srcs, 0.0, 0.5, T, pml_thickness=L, sigma_max=smax)
            gate = 1e-10
        elif kind == "jacobi":
            p.set_diffusivity(0.05 * D)  # strongly diagonally dominant: the iteration converges
            p.solve(dt=0.5, t_total=T, mode="explicit")
            want, taxis = O.solve_fixed(mesh, 0.05 * D, init, lbc, rbc, sidx, srcs, 0.0, 0.5, T, mode="explicit")
            gate = 1e-8
        else:
            kw = dict(tol=float(rng.choice([1e-3, 1e-2])), max_dt=float(rng.choice([0.5, 2.0])))
            p.solve(optimizer=True, dt_init=0.05, t_total=T, adaptive_on_device=False, **kw)
            want, taxis, _ = O.solve_adaptive(mesh, D, init, lbc, rbc, sidx, srcs, 0.0, 0.05, T, **kw)
            gate = 1e-9
    assert p.taxis.shape == taxis.shape, f"{desc}: {p.taxis.shape} vs {taxis.shape}"
    np.testing.assert_allclose(p.taxis, taxis, rtol=1e-9, atol=0, err_msg=desc)
    return G.rel_max(p.snapshot, want), gate, desc

"""Parameter sweeps over ensembles of diffusion models.

NEW capability: ``fiberis.simulator.optimizer`` holds only the adaptive time-step controller
(``tso.py``); the closest thing to a sweep in the reference is the scripted MOOSE pipeline with its
``misfit_calculator`` (src/fiberis/moose/templates/baseline_model_generator.py:418-520).  What is here is
deliberately small: candidate generators on the host, one ensemble solve per batch of candidates on the
GPU(s) with the misfit fused into the run kernel (``PDS1D_Ensemble``), and a shrinking-box random search on
top.  The objective is whatever the ensemble's observations define: with
``set_observations(..., taxis=...)`` / ``set_observed_data`` it is the reference's ``misfit_calculator``
(observed channels on their own time axis, baseline sample, window, channel weights — pinned by
tests/golden/misfit_reference_320x96.npz); candidates whose misfit is NaN (observed times outside the
simulated range) sort last.  Candidates are per-zone log10 diffusivities ``theta`` (M, n_zones); D = 10 ** theta.
"""
from __future__ import annotations

import numpy as np


def grid_candidates(axes):
    """Full tensor grid: ``axes`` = one (lo, hi, n) triple per zone -> theta (prod n, n_zones),
    first zone varying slowest."""
    lines = [np.linspace(lo, hi, int(n)) for lo, hi, n in axes]
    mesh = np.meshgrid(*lines, indexing="ij")
    return np.stack([m.reshape(-1) for m in mesh], axis=1)


def random_candidates(n_models, lo, hi, seed=0):
    """Uniform candidates in the box [lo, hi] (arrays of length n_zones)."""
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, (int(n_models), len(lo)))


def run_sweep(ensemble, theta, zone_of_cell, dt, t_total=None, **solve_kwargs):
    """Evaluate every candidate: sets ``D_m(x) = 10 ** theta[m, zone_of_cell[x]]`` on ``ensemble``
    (mesh, sources, boundary conditions, initial state and observations already set) and solves.
    Under an initialised ``torch.distributed`` group members are sharded over the ranks and the
    misfits all-gathered.  Returns dict(misfit (M,), order (M,) ascending, best, best_theta, best_misfit)."""
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    if ensemble.obs_idx is None:
        raise ValueError("Set the observations first (set_observations).")
    ensemble.set_zonal_diffusivity(10.0 ** theta, zone_of_cell)
    misfit = np.asarray(ensemble.solve(dt=dt, t_total=t_total, **solve_kwargs))
    order = np.argsort(misfit, kind="stable")
    best = int(order[0])
    return dict(misfit=misfit, order=order, best=best, best_theta=theta[best].copy(), best_misfit=float(misfit[best]))


def refine_sweep(ensemble, zone_of_cell, lo, hi, dt, t_total=None, n_models=1024, rounds=4, shrink=0.5, seed=0,
                 **solve_kwargs):
    """Shrinking-box random search: every round draws ``n_models`` candidates in the current box
    (plus the incumbent), evaluates them in one ensemble solve, re-centres the box on the best
    candidate and scales its width by ``shrink`` (clipped to the original bounds).  Returns
    dict(best_theta, best_misfit, history=[(best_misfit, box_lo, box_hi) per round])."""
    lo0, hi0 = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    lo_r, hi_r = lo0.copy(), hi0.copy()
    best_theta, best_misfit, history = None, np.inf, []
    for r in range(int(rounds)):
        theta = random_candidates(n_models, lo_r, hi_r, seed=seed + r)
        if best_theta is not None:
            theta[0] = best_theta  # keep the incumbent: the best misfit never gets worse
        out = run_sweep(ensemble, theta, zone_of_cell, dt, t_total, **solve_kwargs)
        if out["best_misfit"] <= best_misfit:
            best_theta, best_misfit = out["best_theta"], out["best_misfit"]
        history.append((best_misfit, lo_r.copy(), hi_r.copy()))
        half = 0.5 * (hi_r - lo_r) * shrink
        lo_r = np.clip(best_theta - half, lo0, hi0)
        hi_r = np.clip(best_theta + half, lo0, hi0)
    return dict(best_theta=best_theta, best_misfit=best_misfit, history=history)

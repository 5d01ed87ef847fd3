"""CPU oracle (NumPy) for the fibeRIS 1D pressure-diffusion hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module, and only as the
checker / the timed CPU baseline.  ``fiberis_b200`` never imports it: the product path is CUDA or it
fails loudly.

What it restates (reference paths are relative to /root/reference/):

* coefficient assembly .......... ``src/fiberis/simulator/solver/matbuilder.py:21-28,39-81,103-165``
* PML ramp ...................... ``src/fiberis/simulator/solver/matbuilder.py:170-196``
* linear solve .................. ``src/fiberis/simulator/solver/PDESolver_IMP.py:27-28``
  (``scipy.linalg.solve`` -> LAPACK; the arithmetic lives in SciPy/LAPACK, pinned
  ``scipy 1.13.1`` in ``poetry.lock:898-899``, declared ``>=1.8,<2.0`` in ``pyproject.toml:46``.
  SciPy >= 1.15 detects the tridiagonal structure and calls ``?gtsv``; older versions call
  ``?getrf/?getrs``.  Both are Gaussian elimination with partial pivoting and differ by <= 4e-16 per
  solve.  The oracle calls LAPACK ``dgttrf`` once + ``dgttrs`` per step — the same pivoted
  elimination as ``dgtsv`` — or, in ``dense=True`` mode, literally ``scipy.linalg.solve`` on the
  dense matrix as the reference does.)
* Jacobi "explicit" solve ....... ``src/fiberis/simulator/solver/PDESolver_EXP.py:5-39``
* time loop, fixed dt ........... ``src/fiberis/simulator/core/pds.py:243-310,339-340``
* time loop, adaptive dt ........ ``src/fiberis/simulator/core/pds.py:312-336`` +
                                  ``src/fiberis/simulator/optimizer/tso.py:4-50``
* source sampling ............... ``src/fiberis/analyzer/Data1D/core1D.py:245-247`` (``np.interp``,
  clamped at both ends; arithmetic lives in NumPy, pinned 1.26.4 in ``poetry.lock:606-607``)

Parity pinning: ``tests/test_oracle_golden.py`` checks every function here against (1) the
reference's own known-answer tests (``tests/test_simulator_core.py:181-275``,
``tests/test_simulator_solver.py:53-77,112-127,205-254``, ``tests/test_simulator_optimizer.py:28-103``)
and (2) the fixtures in ``tests/golden/`` that ``oracle/gen_golden.py`` produced by running the
unmodified reference in the build container (NumPy 2.3.5 / SciPy 1.18.1).

The ensemble misfit on the step times (``ensemble_misfit``) has NO counterpart in ``fiberis.simulator``: **parity
unpinned** for that one function; its per-member pressures are pinned by looping the reference.  The misfit with
observations on their own time axis (``sampled_misfit``) restates
``src/fiberis/moose/templates/baseline_model_generator.py:479-490`` and is pinned by
``tests/golden/misfit_reference_320x96.npz``, produced by the reference's unmodified ``misfit_calculator``.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import lapack as _lapack
from scipy.linalg import solve as _dense_solve

BC_NONE, BC_DIRICHLET, BC_NEUMANN = 0, 1, 2
_BC_CODE = {"None": BC_NONE, "Dirichlet": BC_DIRICHLET, "Neumann": BC_NEUMANN}


def bc_code(name: str) -> int:
    return _BC_CODE[name]


# ----------------------------------------------------------------------------------------------
# assembly  (matbuilder.py:21-28, 39-81 / 103-165)
# ----------------------------------------------------------------------------------------------
def interface_alphas(mesh, diffusivity, dt):
    """alpha_l, alpha_r for interior rows 1..nx-2 (each of length nx-2).  matbuilder.py:21-28."""
    mesh = np.asarray(mesh, dtype=np.float64)
    D = np.asarray(diffusivity, dtype=np.float64)
    dx = np.diff(mesh)
    d_eff = 2 * D[:-1] * D[1:] / (D[:-1] + D[1:])
    dxm = dx[:-1]
    dxp = dx[1:]
    alpha_l = d_eff[:-1] * dt / (dxm * (dxm + dxp) / 2)
    alpha_r = d_eff[1:] * dt / (dxp * (dxm + dxp) / 2)
    return alpha_l, alpha_r


def build_pml_sigma(mesh, pml_thickness, sigma_max):
    """Linear PML ramp.  matbuilder.py:170-196 (vectorised; same per-element arithmetic)."""
    mesh = np.asarray(mesh, dtype=np.float64)
    sigma = np.zeros(len(mesh))
    if pml_thickness == 0:
        return sigma  # (L - d)/L is never evaluated by the reference when nothing is < 0
    dist_left = mesh - mesh[0]
    dist_right = mesh[-1] - mesh
    left = dist_left < pml_thickness
    sigma[left] = sigma_max * ((pml_thickness - dist_left[left]) / pml_thickness)
    right = dist_right < pml_thickness
    sigma[right] = np.maximum(sigma[right], sigma_max * ((pml_thickness - dist_right[right]) / pml_thickness))
    return sigma


def assemble_tridiag(mesh, diffusivity, dt, lbc, rbc, sourceidx, pml_thickness=0.0, sigma_max=2):
    """Three diagonals (each length nx; dl[0] = du[nx-1] = 0) of the reference's dense A.

    Order of row overwrites follows matbuilder.py:39-81: interior stencil -> boundary rows ->
    source rows (which win over boundary rows) -> PML added to EVERY diagonal entry.
    """
    mesh = np.asarray(mesh, dtype=np.float64)
    nx = len(mesh)
    dl = np.zeros(nx)
    d = np.zeros(nx)
    du = np.zeros(nx)
    if nx > 2:
        al, ar = interface_alphas(mesh, diffusivity, dt)
        dl[1:-1] = -al
        d[1:-1] = 1 + al + ar
        du[1:-1] = -ar
    if lbc == "Dirichlet":
        d[0] = 1
    elif lbc == "Neumann":
        d[0] = -1
        du[0] = 1
    if rbc == "Dirichlet":
        d[-1] = 1
    elif rbc == "Neumann":
        d[-1] = -1
        dl[-1] = 1
    for s in np.atleast_1d(np.asarray(sourceidx)).astype(np.int64):
        dl[s] = 0
        du[s] = 0
        d[s] = 1
    sigma = build_pml_sigma(mesh, pml_thickness, sigma_max)
    d = d + dt * sigma
    return dl, d, du


def dense_from_tridiag(dl, d, du):
    nx = len(d)
    A = np.zeros((nx, nx))
    idx = np.arange(nx)
    A[idx, idx] = d
    A[idx[1:], idx[:-1]] = dl[1:]
    A[idx[:-1], idx[1:]] = du[:-1]
    return A


def assemble_rhs(p_prev, lbc, rbc, sourceidx, source_values):
    """b of matbuilder.py:45-76: copy of the previous state, BC entries zeroed, source entries set
    in list order (later duplicates win)."""
    b = np.array(p_prev, dtype=np.float64, copy=True)
    if lbc in ("Dirichlet", "Neumann"):
        b[0] = 0
    if rbc in ("Dirichlet", "Neumann"):
        b[-1] = 0
    for s, v in zip(np.atleast_1d(np.asarray(sourceidx)).astype(np.int64), np.atleast_1d(source_values)):
        b[s] = v
    return b


# ----------------------------------------------------------------------------------------------
# source sampling  (core1D.py:245-247) and the time axis  (pds.py:282-287,309)
# ----------------------------------------------------------------------------------------------
def source_value(src_taxis, src_data, t):
    return float(np.interp(float(t), src_taxis, src_data))


def time_axis(t0, dt, t_total):
    """The exact sequence ``taxis`` that ``while taxis[-1] < t_total: taxis.append(taxis[-1]+dt)``
    produces (fp64 accumulation; e.g. dt=0.1, t_total=1.0 gives 11 steps)."""
    taxis = [t0]
    while taxis[-1] < t_total:
        taxis.append(taxis[-1] + dt)
    return taxis


# ----------------------------------------------------------------------------------------------
# linear solves
# ----------------------------------------------------------------------------------------------
class _Factor:
    """LAPACK dgttrf factors (partial pivoting) of one tridiagonal matrix."""

    def __init__(self, dl, d, du):
        self.n = len(d)
        if self.n < 3:  # f2py's dgttrf wrapper rejects n < 3 (du2 has n-2 entries): dense dgesv
            self.A = dense_from_tridiag(np.asarray(dl, float), np.asarray(d, float), np.asarray(du, float))
            if np.linalg.det(self.A) == 0:
                raise np.linalg.LinAlgError("singular matrix")
            return
        self.f = _lapack.dgttrf(dl[1:].copy(), d.copy(), du[:-1].copy())
        if self.f[-1] > 0:
            raise np.linalg.LinAlgError("singular matrix")

    def solve(self, b):
        if self.n < 3:
            return np.linalg.solve(self.A, b)
        dl_, d_, du_, du2, ipiv, _ = self.f
        x, info = _lapack.dgttrs(dl_, d_, du_, du2, ipiv, b)
        assert info == 0
        return x


def solve_tridiag(dl, d, du, b):
    return _Factor(dl, d, du).solve(np.asarray(b, dtype=np.float64))


def jacobi(dl, d, du, b, tol=1e-10, max_iter=1000):
    """PDESolver_EXP.py:5-39 on a tridiagonal matrix: Jacobi from x=0, inf-norm stop, silent on
    non-convergence (the last iterate is returned, as the reference does)."""
    x = np.zeros_like(b, dtype=np.float64)
    for _ in range(max_iter):
        off = np.zeros_like(x)
        off[1:] += dl[1:] * x[:-1]
        off[:-1] += du[:-1] * x[1:]
        x_new = (b - off) / d
        if np.linalg.norm(x_new - x, ord=np.inf) < tol:
            return x_new
        x = x_new
    return x


# ----------------------------------------------------------------------------------------------
# time loops
# ----------------------------------------------------------------------------------------------
def _src_table(sources, taxis, t0):
    """sources: list of (src_taxis, src_data).  Returns (len(taxis), n_src): value of source k
    at taxis[n] - t0, i.e. the value the step STARTING at taxis[n] uses (matbuilder.py:73)."""
    tt = np.asarray(taxis, dtype=np.float64) - t0
    return np.stack([np.interp(tt, st, sd) for st, sd in sources], axis=1)


def solve_fixed(mesh, diffusivity, initial, lbc, rbc, sourceidx, sources, t0, dt, t_total,
                mode="implicit", dense=False, pml_thickness=0.0, sigma_max=2, rebuild_every_step=False):
    """pds.py:243-310 for optimizer=False.  Returns (snapshot (nt+1,nx), taxis (nt+1,)).

    ``dense=True`` solves the dense (nx, nx) system with ``scipy.linalg.solve`` as the reference
    does; with ``rebuild_every_step=True`` the dense matrix is also re-assembled before every solve
    (pds.py:294), which is the reference's cost profile minus its Python ``for`` loops — the mode
    ``bench.py --impl reference`` times."""
    taxis = time_axis(t0, dt, t_total)
    sidx = np.atleast_1d(np.asarray(sourceidx)).astype(np.int64)
    S = _src_table(sources, taxis, t0)
    dl, d, du = assemble_tridiag(mesh, diffusivity, dt, lbc, rbc, sidx, pml_thickness, sigma_max)
    snap = np.empty((len(taxis), len(mesh)))
    snap[0] = initial
    if mode == "implicit" and not dense:
        fac = _Factor(dl, d, du)
    A = dense_from_tridiag(dl, d, du) if dense else None
    for n in range(len(taxis) - 1):
        b = assemble_rhs(snap[n], lbc, rbc, sidx, S[n])
        if mode != "implicit":
            snap[n + 1] = jacobi(dl, d, du, b)
        elif dense:
            if rebuild_every_step:
                A = dense_from_tridiag(*assemble_tridiag(mesh, diffusivity, dt, lbc, rbc, sidx, pml_thickness,
                                                         sigma_max))
            snap[n + 1] = _dense_solve(A, b)
        else:
            snap[n + 1] = fac.solve(b)
    return snap, np.array(taxis)


def adjust_dt(dt, error, safety_factor=0.9, p=2, max_dt=40, min_dt=1e-4, tol=1e-3):
    """tso.py:35-50."""
    if error < 1e-14:
        ratio = 1.0
    else:
        ratio = (tol / error) ** (1.0 / p)
    dt_new = dt * safety_factor * ratio
    return max(min_dt, min(dt_new, max_dt))


def solve_adaptive(mesh, diffusivity, initial, lbc, rbc, sourceidx, sources, t0, dt_init, t_total,
                   tol=1e-3, max_iter=None, mode="implicit", **adjust_kwargs):
    """pds.py:312-336 + tso.py:4-33 for optimizer=True.

    Quirks preserved: the user's ``tol`` is used for accept/reject only and is NOT forwarded to
    ``adjust_dt`` (tso.py:4,21 — it binds to the positional parameter); both half steps use the
    source value at the START of the step (taxis is not advanced, matbuilder.py:73); the accepted
    state is the FULL-step solution (tso.py:26).  ``max_iter`` bounds the loop (the reference can
    spin forever when err is nan).  ``mode`` other than ``'implicit'``: the FULL step of every iteration is
    solved with Jacobi, the two half steps stay implicit (pds.py:296-300 against :319,325).
    Returns (snapshot, taxis, n_iterations).
    """
    sidx = np.atleast_1d(np.asarray(sourceidx)).astype(np.int64)
    adjust_kwargs = {k: v for k, v in adjust_kwargs.items()
                     if k in ("safety_factor", "p", "max_dt", "min_dt")}
    snaps = [np.asarray(initial, dtype=np.float64)]
    taxis = [t0]
    dt = dt_init
    it = 0
    while taxis[-1] < t_total:
        if max_iter is not None and it >= max_iter:
            break
        it += 1
        sv = [np.interp(float(taxis[-1] - t0), st, sd) for st, sd in sources]
        b = assemble_rhs(snaps[-1], lbc, rbc, sidx, sv)
        if mode == "implicit":
            full = solve_tridiag(*assemble_tridiag(mesh, diffusivity, dt, lbc, rbc, sidx), b)
        else:
            full = jacobi(*assemble_tridiag(mesh, diffusivity, dt, lbc, rbc, sidx), b)
        fh = _Factor(*assemble_tridiag(mesh, diffusivity, dt / 2, lbc, rbc, sidx))
        mid = fh.solve(b)
        half = fh.solve(assemble_rhs(mid, lbc, rbc, sidx, sv))
        with np.errstate(invalid="ignore", divide="ignore"):
            err = np.linalg.norm(full - half) / np.linalg.norm(full)
        dt_next = adjust_dt(dt, err, **adjust_kwargs)
        if err <= tol:
            snaps.append(full)
            taxis.append(taxis[-1] + dt)
        dt = dt_next
    return np.array(snaps), np.array(taxis), it


# ----------------------------------------------------------------------------------------------
# ensembles (new capability; per-member pressures = the reference looped over members)
# ----------------------------------------------------------------------------------------------
def zonal_diffusivity(theta, zone_of_cell):
    """D_m(x) = 10 ** theta[m, zone(x)]  (SURVEY.md section 8(d), configs C3/C4)."""
    theta = np.asarray(theta, dtype=np.float64)
    return np.power(10.0, theta)[..., np.asarray(zone_of_cell)]


def ensemble_misfit(snapshot, obs_idx, obs, weights=None):
    """misfit = sum_{n=1..nt} sum_k w_k (p^n[obs_idx_k] - obs[n, k])^2 .

    PARITY UNPINNED: fiberis.simulator defines no misfit.  Modelled on the SSE in
    ``src/fiberis/moose/templates/baseline_model_generator.py:394-404`` (channel-weighted sum of
    squared residuals), without the interpolation/baseline-subtraction steps because observation
    times coincide with the step times here.  ``obs`` has one row per entry of ``taxis``; row 0
    (the initial state) does not contribute.
    """
    snapshot = np.asarray(snapshot)
    obs = np.asarray(obs)
    w = np.ones(len(obs_idx)) if weights is None else np.asarray(weights, dtype=np.float64)
    r = snapshot[1:, obs_idx] - obs[1:]
    return float(np.sum(w[None, :] * r * r))


def sampled_misfit(snapshot, taxis, obs_idx, obs, obs_taxis, weights=None, baseline_index=1, window=(1, -1),
                   time_offset=0.0, per_channel=False):
    """The reference's history-matching misfit for observations on their OWN time axis
    (``src/fiberis/moose/templates/baseline_model_generator.py:479-490``):

    * the simulated channel is interpolated onto the observed times — ``Data1D.interpolate``
      (``src/fiberis/analyzer/Data1D/core1D.py:786-812``): points ``obs_taxis + (obs_start - sim_start)`` in the simulated
      frame, ``np.interp(..., left=nan, right=nan)``;
    * "initial time calibration": the sample at ``baseline_index`` (the reference: 1) is subtracted (``:483``);
    * ``sum((sim[first:last] - obs[first:last]) ** 2)`` with the reference's window ``[1:-1]`` (``:486``), times the
      channel weight (``:488``), summed over channels.

    ``snapshot (n_t, nx)``, ``taxis (n_t)``; ``obs (n_samples, n_channels)``; ``time_offset`` = observed start_time minus
    simulated start_time in seconds."""
    snapshot = np.asarray(snapshot, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64)
    w = np.ones(len(obs_idx)) if weights is None else np.asarray(weights, dtype=np.float64)
    points = np.asarray(obs_taxis, dtype=np.float64) + time_offset
    sl = slice(*window) if window is not None else slice(None)
    out = []
    for k, cell in enumerate(obs_idx):
        sim = np.interp(points, taxis, snapshot[:, cell], left=np.nan, right=np.nan)
        if baseline_index is not None:
            sim = sim - sim[baseline_index]
        out.append(w[k] * np.sum((sim[sl] - obs[sl, k]) ** 2))
    return np.asarray(out) if per_channel else float(np.sum(out))

"""Adaptive time-step controller (step doubling), API of
/root/reference/src/fiberis/simulator/optimizer/tso.py:4-50.

``adjust_dt`` is scalar host control logic.  ``time_sampling_optimizer`` measures the step-doubling
error ``||full - half||_2 / ||full||_2`` with the CUDA reduction kernel K5
(``fbr_error_norms_f64``); ``PDS1D_*.solve(optimizer=True)`` calls the same kernel directly on
device-resident states through ``step_error`` and only moves the scalar to the host.
"""
import numpy as np


def adjust_dt(dt, error, **kwargs):
    """Next dt: clamp(dt * safety_factor * (tol/error)**(1/p), min_dt, max_dt).

    NB (reference behaviour, kept): ``tol`` here comes from ``kwargs`` only — a ``tol`` passed to
    ``time_sampling_optimizer`` binds to that function's own parameter and is not forwarded.
    """
    safety_factor = kwargs.get("safety_factor", 0.9)
    p = kwargs.get("p", 2)
    max_dt = kwargs.get("max_dt", 40)
    min_dt = kwargs.get("min_dt", 1e-4)
    tol = kwargs.get("tol", 1e-3)
    ratio = 1.0 if error < 1e-14 else (tol / error) ** (1.0 / p)
    return max(min_dt, min(dt * safety_factor * ratio, max_dt))


def step_error(full_dev, half_dev) -> float:
    """err of one model from two device-resident (1, nx) states (one 8-byte D2H copy)."""
    from fiberis_b200 import _engine
    return float(_engine.error_norms(full_dev, half_dev)[0].item())


def time_sampling_optimizer(pds1d, full_step_solution, half_step_solution, dt, tol=1e-3, *args, **kwargs) -> float:
    """Accept (append ``full_step_solution`` and ``taxis[-1] + dt``) when err <= tol; return next dt."""
    from fiberis_b200 import _engine
    full = np.asarray(full_step_solution, dtype=np.float64)
    half = np.asarray(half_step_solution, dtype=np.float64)
    error = step_error(_engine.to_device(full[None, :]), _engine.to_device(half[None, :]))
    return decide(pds1d, full_step_solution, error, dt, tol, **kwargs)


def decide(pds1d, full_step_solution, error, dt, tol=1e-3, **kwargs) -> float:
    """The accept / reject branch of tso.py:21-33 for an already measured ``error``."""
    updated_dt = adjust_dt(dt, error, **kwargs)
    if error <= tol:
        pds1d.snapshot.append(full_step_solution)
        pds1d.taxis.append(pds1d.taxis[-1] + dt)
        pds1d.record_log("Dynamic time sampling updated. dt = ", dt, "Error = ", error, "Next loop dt will be",
                         updated_dt, "\n")
    else:
        pds1d.record_log("Dynamic time sampling rejected. dt = ", dt, "Next loop dt will be", updated_dt, "\n")
    return updated_dt

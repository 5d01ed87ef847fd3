"""Implicit linear solve, API of
/root/reference/src/fiberis/simulator/solver/PDESolver_IMP.py:9-34.

``A`` arrives as the dense matrix the reference's ``matbuilder`` produces.  A tridiagonal ``A`` (the
only kind the simulator builds) is solved by the CUDA batched tridiagonal kernel K2
(``fbr_solve_tridiag_f64``: register-resident elimination without pivoting; a zero pivot of a regular
matrix falls through to the pivoted variant — LAPACK's DGTSV algorithm, what the reference's
``scipy.linalg.solve`` does).  Any other dense ``A`` goes to ``fbr_solve_dense_f64`` (Gaussian elimination
with partial pivoting by one CTA) — no library solve is left behind this API.
All three ``solver`` names of the reference select the same device path; an unknown name raises ``ValueError('Invalid solver type.')`` like the reference.
"""
import numpy as np

from fiberis_b200 import _engine


def split_tridiagonal(A):
    """(dl, d, du) if A is tridiagonal, else None."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    if A.ndim != 2 or A.shape[1] != n:
        raise ValueError("A must be a square matrix.")
    band = np.abs(np.subtract.outer(np.arange(n), np.arange(n))) <= 1
    if np.any(A[~band] != 0):
        return None
    dl, du = np.zeros(n), np.zeros(n)
    dl[1:] = np.diagonal(A, -1)
    du[:-1] = np.diagonal(A, 1)
    return dl, np.diagonal(A).copy(), du


def solver_implicit(A, b, **kwargs):
    solver = kwargs.get("solver", "scipy")
    if solver not in ("scipy", "numpy", "direct"):
        raise ValueError("Invalid solver type.")
    b = np.asarray(b, dtype=np.float64)
    tri = split_tridiagonal(A)

    if tri is None or len(b) < 2:
        return _engine.solve_dense(_engine.to_device(np.asarray(A, dtype=np.float64)), _engine.to_device(b)).cpu().numpy()
    dl, d, du = (_engine.to_device(v[None, :]) for v in tri)
    return _engine.solve_tridiag(dl, d, du, _engine.to_device(b[None, :]))[0].cpu().numpy()

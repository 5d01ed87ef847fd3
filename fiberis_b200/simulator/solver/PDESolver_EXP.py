""""Explicit" solver = Jacobi iteration on ``A x = b``, API of
/root/reference/src/fiberis/simulator/solver/PDESolver_EXP.py:5-39.

Tridiagonal ``A`` runs in the CUDA Jacobi kernel (``fbr_jacobi_f64``: one CTA per system, iterate
in shared memory); any other dense matrix in ``fbr_jacobi_dense_f64`` (one CTA, a warp per row).  As in the reference the
function prints how it ended and returns the last iterate even without convergence.
"""
import numpy as np

from fiberis_b200 import _engine
from fiberis_b200.simulator.solver.PDESolver_IMP import split_tridiagonal


def solver_explicit(A, b, tol=1e-10, max_iter=1000, **kwargs):
    b = np.asarray(b, dtype=np.float64)
    tri = split_tridiagonal(A)
    if tri is not None and len(b) <= 8192:
        dl, d, du = (_engine.to_device(v[None, :]) for v in tri)
        x, iters = _engine.jacobi(dl, d, du, _engine.to_device(b[None, :]), tol, max_iter)
        used = int(iters[0].item())
        x = x[0].cpu().numpy()
    else:  # any other square matrix: the dense Jacobi kernel (one CTA, a warp per row)
        A_d = _engine.to_device(np.ascontiguousarray(np.asarray(A, dtype=np.float64)))
        x_d, used = _engine.jacobi_dense(A_d, _engine.to_device(b), tol, max_iter)
        x = x_d.cpu().numpy()
    if used > 0:
        print(f"Converged in {used} iterations.")
    else:
        print("Warning: Maximum iterations reached without convergence.")
    return x

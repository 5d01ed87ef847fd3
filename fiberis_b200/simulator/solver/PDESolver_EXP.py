""""Explicit" solver = Jacobi iteration on ``A x = b``, API of
/root/reference/src/fiberis/simulator/solver/PDESolver_EXP.py:5-39.

Tridiagonal ``A`` runs in the CUDA Jacobi kernel (``fbr_jacobi_f64``: one CTA per system, iterate
in shared memory); other dense matrices iterate with device mat-vecs.  As in the reference the
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
    else:
        import torch
        A_d = _engine.to_device(np.asarray(A, dtype=np.float64))
        b_d = _engine.to_device(b)
        diag = torch.diagonal(A_d)
        x_d = torch.zeros_like(b_d)
        used = -max_iter
        for it in range(max_iter):
            x_new = (b_d - (A_d @ x_d - diag * x_d)) / diag
            done = bool((x_new - x_d).abs().max() < tol)
            x_d = x_new
            if done:
                used = it + 1
                break
        x = x_d.cpu().numpy()
    if used > 0:
        print(f"Converged in {used} iterations.")
    else:
        print("Warning: Maximum iterations reached without convergence.")
    return x

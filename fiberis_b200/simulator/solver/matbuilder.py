"""Assembly of the implicit system, API of
/root/reference/src/fiberis/simulator/solver/matbuilder.py:6-196.

The coefficients are produced by the CUDA assembly kernel K1 (``fbr_assemble_f64``) as three
diagonals — bit-identical to the reference's NumPy arithmetic — and, for API parity with the
reference's tests and debugging workflows, scattered into the dense ``(nx, nx)`` matrix these
functions have always returned.  ``PDS1D_*.solve`` does not go through the dense form.
"""
import numpy as np

from fiberis_b200 import _engine


def _source_values(pds1d):
    """Sources are sampled at the START of the step: taxis[-1] - t0 (reference :73, :152)."""
    t = pds1d.taxis[-1] - pds1d.t0
    if isinstance(pds1d.source, list):
        return [s.get_value_by_time(t) for s in pds1d.source]
    return [pds1d.source.get_value_by_time(t)]


def tridiagonal_system(pds1d, dt, pml_thickness, sigma_max):
    """(dl, d, du, b) as host arrays, computed on the device (K1 + right-hand-side kernel)."""
    mesh = np.asarray(pds1d.mesh, dtype=np.float64)
    nx = len(mesh)
    sidx = np.atleast_1d(np.asarray(pds1d.sourceidx)).astype(np.int64)
    sidx = np.where(sidx < 0, sidx + nx, sidx)  # NumPy row indexing accepts negative indices
    mesh_d = _engine.to_device(mesh)
    diff_d = _engine.to_device(pds1d.diffusivity)
    sidx_d, n_src = _engine.index_tensor(sidx)
    dl, d, du = _engine.assemble(mesh_d, diff_d, 1, nx, dt, pds1d.lbc, pds1d.rbc, sidx_d, n_src, pml_thickness,
                                 sigma_max)
    prev = _engine.to_device(np.asarray(pds1d.snapshot[-1], dtype=np.float64)[None, :])
    vals = _engine.to_device(np.asarray(_source_values(pds1d), dtype=np.float64))
    b = _engine.rhs(prev, pds1d.lbc, pds1d.rbc, sidx_d, n_src, vals)
    return dl[0].cpu().numpy(), d[0].cpu().numpy(), du[0].cpu().numpy(), b[0].cpu().numpy()


def _dense(dl, d, du):
    nx = len(d)
    A = np.zeros((nx, nx))
    k = np.arange(nx)
    A[k, k] = d
    A[k[1:], k[:-1]] = dl[1:]
    A[k[:-1], k[1:]] = du[:-1]
    return A


def matrix_builder_1d_single_source(pds1d, dt, pml_thickness=0.0, sigma_max=2):
    dl, d, du, b = tridiagonal_system(pds1d, dt, pml_thickness, sigma_max)
    return _dense(dl, d, du), b


def matrix_builder_1d_multi_source(pds1d, dt, pml_thickness=0.0, sigma_max=10):
    dl, d, du, b = tridiagonal_system(pds1d, dt, pml_thickness, sigma_max)
    return _dense(dl, d, du), b


def build_pml_sigma(mesh, pml_thickness, sigma_max):
    """Linear PML ramp towards both ends of the mesh (device kernel ``fbr_pml_sigma_f64``)."""
    mesh_d = _engine.to_device(np.asarray(mesh, dtype=np.float64))
    return _engine.pml_sigma(mesh_d, pml_thickness, sigma_max).cpu().numpy()

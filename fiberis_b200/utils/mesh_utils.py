"""Mesh helpers users build non-uniform meshes and source indices with (host-side, NumPy).

Same behaviour as /root/reference/src/fiberis/utils/mesh_utils.py:10-80.
"""
from __future__ import annotations

import numpy as np


def refine_mesh(x, refine_range, refine_factor):
    """Replace the nodes inside ``refine_range`` by ``refine_factor`` times as many, evenly spaced."""
    if not isinstance(x, np.ndarray):
        raise ValueError("Input x must be a numpy.ndarray representing a 1D array.")
    if len(refine_range) != 2 or refine_range[0] >= refine_range[1]:
        raise ValueError("refine_range must be a tuple of length 2 with start < end.")
    if refine_factor < 1:
        raise ValueError("refine_factor must be greater than or equal to 1.")
    lo, hi = refine_range
    first = np.searchsorted(x, lo, side="left")
    last = np.searchsorted(x, hi, side="right")
    fine = np.linspace(lo, hi, (last - first) * refine_factor + 1)
    return np.sort(np.concatenate((x[:first], x[last:], fine)))


def locate(x, frac_loc):
    """(index, value) of the node of ``x`` nearest to ``frac_loc``."""
    if not isinstance(x, np.ndarray):
        raise ValueError("Input x must be a numpy.ndarray.")
    if x.size == 0:
        raise ValueError("Input x cannot be an empty array.")
    if not np.isscalar(frac_loc):
        raise ValueError("frac_loc must be a scalar value.")
    ind = np.argmin(np.abs(x - frac_loc))
    return ind, x[ind]

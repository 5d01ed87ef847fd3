"""Helpers shared by the parity tests: load a fixture written by oracle/gen_golden.py."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SOLVE_CASES = [
    "c1_single_200x1000", "multi_nonuniform_96x80", "single_t0_default_ttotal_33",
    "multi_boundary_sources_none_bc_40", "multi_duplicate_sources_24", "single_tiny_2", "single_tiny_3",
    "c3_member_1024x60", "single_stiff_300x20",
]
ADAPTIVE_CASES = ["adaptive_single_50", "adaptive_multi_50_kwargs"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    for k in ("lbc", "rbc"):
        if k in g:
            g[k] = str(g[k])
    if "src_taxis" in g and g["src_taxis"].ndim == 2:
        srcs = []
        for t, d in zip(g["src_taxis"], g["src_data"]):
            keep = ~np.isnan(t)
            srcs.append((t[keep], d[keep]))
        g["sources"] = srcs
    for k in ("t0", "dt", "t_total", "dt_init", "tol", "safety_factor", "max_dt", "min_dt", "p"):
        if k in g:
            g[k] = float(g[k])
    if "multi" in g:
        g["multi"] = bool(g["multi"])
    return g


def rel_max(a, b):
    """max|a-b| / max|b|  — the parity gate of SURVEY.md section 8(d)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))

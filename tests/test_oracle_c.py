"""Pins the C oracle (oracle/pds_oracle.c) against the NumPy oracle, which is itself pinned against
the reference (tests/test_oracle_golden.py).  CPU only."""
import numpy as np
import numpy.testing as npt
import pytest

from oracle import c_oracle as CO
from oracle import pds_oracle as O
from tests import _golden as G


@pytest.mark.parametrize("name", G.SOLVE_CASES)
def test_c_oracle_vs_reference_fixture(name):
    g = G.load(name)
    n_steps = len(g["taxis"]) - 1
    table = O._src_table(g["sources"], g["taxis"][:-1], g["t0"])
    snap, final, _ = CO.run(g["mesh"], g["diffusivity"], g["initial"], n_steps, g["dt"], g["lbc"], g["rbc"],
                            g["sourceidx"], table)
    full = np.concatenate([g["initial"][None, :], snap])
    assert G.rel_max(full[g["rows"]], g["snapshot_rows"]) <= 1e-13
    npt.assert_array_equal(final, snap[-1])


def test_c_assembly_bit_exact():
    z = G.load("matbuilder_pml_21")
    for key, (lbc, rbc, sidx, L, smax) in {"1": ("Neumann", "Dirichlet", [10], 4.0, 2.0),
                                           "2": ("Neumann", "Dirichlet", [10], 3.0, 5.0),
                                           "3": ("Dirichlet", "Neumann", [3, 17], 4.0, 10.0)}.items():
        dl, d, du = CO.assemble(z["mesh"], z["diffusivity"], 0.3, lbc, rbc, sidx, L, smax)
        npt.assert_array_equal(O.dense_from_tridiag(dl, d, du), z["A" + key])


def test_c_ensemble_matches_numpy_oracle():
    rng = np.random.default_rng(5)
    nx, M, n_steps = 64, 6, 40
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    zone = (np.arange(nx) * 4) // nx
    zv = 10 ** rng.uniform(-1, 1, (M, 4))
    srcs = [(np.linspace(0, 40, 9), rng.uniform(-5, 5, 9)) for _ in range(2)]
    sidx = [16, 48]
    taxis = O.time_axis(0.0, 1.0, float(n_steps))
    table = O._src_table(srcs, taxis[:-1], 0.0)
    obs_idx = np.array([3, 30, 60])
    ref0, _ = O.solve_fixed(mesh, zv[0][zone], np.zeros(nx), "Neumann", "Neumann", sidx, srcs, 0.0, 1.0, float(n_steps))
    obs = ref0[:, obs_idx]
    w = np.array([1.0, 0.5, 2.0])
    mis, final = CO.run_ensemble(mesh, zv, zone, np.zeros(nx), n_steps, 1.0, "Neumann", "Neumann", sidx, table,
                                 obs_idx, obs, w, want_final=True, threads=2)
    for m in range(M):
        snap, _ = O.solve_fixed(mesh, zv[m][zone], np.zeros(nx), "Neumann", "Neumann", sidx, srcs, 0.0, 1.0,
                                float(n_steps))
        assert G.rel_max(final[m], snap[-1]) <= 1e-13
        assert mis[m] == pytest.approx(O.ensemble_misfit(snap, obs_idx, obs, w), rel=1e-11, abs=1e-22)
    assert mis[0] <= 1e-24

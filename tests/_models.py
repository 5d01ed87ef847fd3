"""Builders shared by the GPU parity tests: a fixture dict -> a configured fiberis_b200 model."""
import datetime

import numpy as np

START = datetime.datetime(2020, 1, 1)


def sources_of(srcs):
    from fiberis_b200.analyzer.Data1D.core1D import Data1D
    return [Data1D(data=np.asarray(d, float), taxis=np.asarray(t, float), start_time=START) for t, d in srcs]


def model(mesh, D, initial, lbc, rbc, sidx, srcs, t0, multi):
    from fiberis_b200.simulator.core import pds
    p = pds.PDS1D_MultiSource() if multi else pds.PDS1D_SingleSource()
    p.set_mesh(np.asarray(mesh, float))
    s = sources_of(srcs)
    p.set_source(s if multi else s[0])
    p.set_bcs(lbc, rbc)
    p.set_initial(np.asarray(initial, float))
    p.set_diffusivity(np.asarray(D, float))
    p.set_sourceidx([int(i) for i in sidx] if multi else int(sidx[0]))
    p.set_t0(t0)
    return p


def from_fixture(g):
    return model(g["mesh"], g["diffusivity"], g["initial"], g["lbc"], g["rbc"], g["sourceidx"], g["sources"],
                 g["t0"], g["multi"])


def random_case(seed, nx, n_src=3, T=32.0, knots=12, lbc=None, rbc=None):
    rng = np.random.default_rng(seed)
    mesh = np.cumsum(rng.uniform(0.5, 2.0, nx))
    D = 10 ** rng.uniform(-1, 1, nx)
    init = rng.uniform(0, 1, nx)
    sidx = sorted(rng.choice(np.arange(1, nx - 1), size=min(n_src, nx - 2), replace=False).tolist())
    srcs = [(np.linspace(0, T, knots), rng.uniform(-10, 10, knots)) for _ in sidx]
    lbc = lbc or rng.choice(["Dirichlet", "Neumann"])
    rbc = rbc or rng.choice(["Dirichlet", "Neumann"])
    return dict(mesh=mesh, diffusivity=D, initial=init, lbc=str(lbc), rbc=str(rbc), sourceidx=np.array(sidx),
                sources=srcs, t0=0.0, multi=True)

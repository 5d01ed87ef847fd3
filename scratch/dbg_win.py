import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import _models as Mo
from tests import _golden as G
from oracle import pds_oracle as O
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
c = Mo.random_case(4242 + nx, nx, n_src=8, T=20.0, knots=16, lbc="Neumann", rbc="Dirichlet")
want, taxis = O.solve_fixed(c["mesh"], c["diffusivity"], c["initial"], "Neumann", "Dirichlet", c["sourceidx"], c["sources"], 0.0, 0.5, 20.0)
for rep in range(3):
    p = Mo.from_fixture(c)
    with contextlib.redirect_stdout(io.StringIO()):
        p.solve(dt=0.5, t_total=20.0, long_mesh="windowed")
    err = np.abs(p.snapshot - want).max(axis=1) / np.abs(want).max()
    bad = np.nonzero(err > 1e-10)[0]
    print(dict(os.environ).get("FBR_WIN_GENERAL"), dict(os.environ).get("FBR_WIN_NO_PDL"), p.long_mesh_path, "max err", err.max(), "first bad row", bad[:3],
          "bad cells in that row", np.nonzero(np.abs(p.snapshot[bad[0]] - want[bad[0]]) > 1e-9)[0][[0, -1]] if len(bad) else None)

import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from tests import _golden as G, _models as Mo
g = G.load("c1_single_200x1000")
for env in ({}, {"FBR_K3_X_MIN_NX": "1"}, {"FBR_CPT": "8"}, {"FBR_CPT": "4"}, {"FBR_CPT": "1"}):
    for k in ("FBR_K3_X_MIN_NX", "FBR_CPT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for oc in (True, False):
        p = Mo.from_fixture(g)
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            p.solve(dt=g["dt"], t_total=g["t_total"], dtype="float32", on_chip_assembly=oc)
        print(env, "on_chip", oc, "rel_max %.3e" % G.rel_max(p.snapshot[g["rows"]], g["snapshot_rows"]))

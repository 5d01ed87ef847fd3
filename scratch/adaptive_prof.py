import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_ensemble import make_ensemble
e, c = make_ensemble(5, 592, 1024, 200)
out = e.solve_adaptive(dt_init=0.05, t_total=60.0, tol=1e-3, max_dt=4.0)
print(int(out["iterations"].sum()))

"""C5's mesh through the exact one-pass kernel for a few steps (ncu target): python scratch/k3p_run.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
w = bench.make_single("c5")
w = dict(w, n_steps=n, T=n * w["dt"])
p = bench.build_single(w)
p.solve(dt=w["dt"], t_total=w["T"], snapshot_every=n, long_mesh="exact", _stream_above=0)
print(p.long_mesh_path, float(p.snapshot[-1].sum()))

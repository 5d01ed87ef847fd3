"""cProfile of one end-to-end C1 solve (host timeline of PDS1D_MultiSource.solve).  Scratch tool."""
import cProfile, io, os, pstats, sys, contextlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
w = bench.make_single("c1")
p = bench.build_single(w)
for _ in range(5):
    with contextlib.redirect_stdout(io.StringIO()):
        p.solve(dt=w["dt"], t_total=w["T"])
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        p.solve(dt=w["dt"], t_total=w["T"])
    ts.append(time.perf_counter() - t0)
print("e2e ms: min %.3f median %.3f" % (min(ts) * 1e3, sorted(ts)[10] * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    with contextlib.redirect_stdout(io.StringIO()):
        p.solve(dt=w["dt"], t_total=w["T"])
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])

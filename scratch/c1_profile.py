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
rows = []
for e in pr.getstats():  # raw stats: microseconds per solve() call
    code = e.code
    name = code if isinstance(code, str) else f"{os.path.basename(code.co_filename)}:{code.co_firstlineno}({code.co_name})"
    rows.append((e.inlinetime / 20 * 1e6, e.totaltime / 20 * 1e6, e.callcount / 20, name))
rows.sort(reverse=True)
print(f"{'inline us':>10s} {'total us':>10s} {'calls':>6s}  function   (per solve() call)")
for r in rows[:36]:
    print(f"{r[0]:10.1f} {r[1]:10.1f} {r[2]:6.1f}  {r[3][:100]}")

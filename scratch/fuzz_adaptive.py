"""Randomised parity sweep of the device-side adaptive loop (tests/_fuzz.py).  python scratch/fuzz_adaptive.py [n] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests._fuzz import run_adaptive_case
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst, fails, t0 = 0.0, 0, time.time()
for _ in range(n_cases):
    try:
        err, desc = run_adaptive_case(rng)
        worst = max(worst, err)
        assert err <= 1e-9, f"{desc}: {err:.2e}"
    except Exception as ex:
        fails += 1
        print("FAIL", type(ex).__name__, str(ex)[:400])
print(f"{n_cases} cases, {fails} failures, worst rel err {worst:.2e}, {time.time() - t0:.0f} s")

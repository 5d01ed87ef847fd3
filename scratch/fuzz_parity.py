"""Randomised parity sweep against the NumPy oracle over every fixed-dt path (tests/_fuzz.py draws the cases).
Scratch tool: python scratch/fuzz_parity.py [n] [seed] [float64|float32]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests._fuzz import run_case

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dtype = sys.argv[3] if len(sys.argv) > 3 else "float64"
gate = 1e-10 if dtype == "float64" else 1e-5
worst, fails, t_start = 0.0, 0, time.time()
for case in range(n_cases):
    try:
        err, desc = run_case(rng, dtype)
        worst = max(worst, err)
        assert err <= gate, f"{desc}: rel err {err:.2e}"
    except Exception as ex:  # keep going: list every failing configuration
        fails += 1
        print("FAIL", type(ex).__name__, str(ex)[:300])
print(f"{n_cases} cases, {fails} failures, worst rel err {worst:.2e}, {time.time() - t_start:.0f} s")

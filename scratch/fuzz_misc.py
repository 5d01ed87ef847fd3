import sys, time
sys.path.insert(0, ".")
import numpy as np
from tests._fuzz import run_misc_case
rng = np.random.default_rng(3); fails = 0; worst = {}
t0 = time.time()
for _ in range(300):
    try:
        err, gate, desc = run_misc_case(rng)
        k = desc.split()[0]; worst[k] = max(worst.get(k, 0), err)
        assert err <= gate, f"{desc}: {err:.2e}"
    except Exception as ex:
        fails += 1; print("FAIL", type(ex).__name__, str(ex)[:300])
print(fails, "failures", worst, round(time.time() - t0), "s")

"""Randomised parity: 150 configurations drawn with a fixed seed over every fixed-dt path (tests/_fuzz.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_configurations_match_the_oracle(seed):
    from tests._fuzz import run_case
    rng = np.random.default_rng(seed)
    for _ in range(50):
        err, desc = run_case(rng)
        assert err <= 1e-10, f"{desc}: {err:.2e}"


def test_random_adaptive_runs_match_the_oracle():
    from tests._fuzz import run_adaptive_case
    rng = np.random.default_rng(21)
    for _ in range(12):
        err, desc = run_adaptive_case(rng)
        assert err <= 1e-9, f"{desc}: {err:.2e}"


def test_random_configurations_fp32_mode():
    """fp32 mode (north_star: stated separately at 1e-5) over random single models and ensembles up to 4096 cells."""
    from tests._fuzz import run_case
    rng = np.random.default_rng(31)
    for _ in range(60):
        err, desc = run_case(rng, "float32")
        assert err <= 1e-5, f"{desc}: {err:.2e}"


def test_random_pml_jacobi_and_host_driven_adaptive_runs():
    from tests._fuzz import run_misc_case
    rng = np.random.default_rng(41)
    for _ in range(24):
        err, gate, desc = run_misc_case(rng)
        assert err <= gate, f"{desc}: {err:.2e}"

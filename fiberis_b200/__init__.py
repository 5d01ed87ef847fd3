"""fiberis_b200 — B200-native (sm_100a) implementation of fibeRIS's 1D pressure-diffusion simulator.

Import layout mirrors the reference so existing scripts only change the top-level package name:

    from fiberis_b200.simulator.core import pds            # PDS1D_SingleSource / PDS1D_MultiSource
    from fiberis_b200.simulator.core.ensemble import PDS1D_Ensemble   (new: batched models)
    from fiberis_b200.simulator.solver import matbuilder, PDESolver_IMP, PDESolver_EXP
    from fiberis_b200.simulator.optimizer import tso
    from fiberis_b200.analyzer.Data1D.core1D import Data1D
    from fiberis_b200.analyzer.Data2D.core2D import Data2D
    from fiberis_b200.utils import mesh_utils

All arithmetic of the path runs in hand-written CUDA behind the C ABI in ``include/fiberis_b200.h``;
there is no CPU fallback.  ``install_as_fiberis()`` aliases these modules under the reference's own
names (``fiberis.simulator...``) for scripts that cannot be edited.
"""
__version__ = "0.1.0"


def install_as_fiberis():
    """Register this package's modules under the reference's import paths (``fiberis.*``)."""
    import importlib
    import sys
    import types

    names = [
        "simulator", "simulator.core", "simulator.core.pds", "simulator.core.bcs", "simulator.core.ensemble",
        "simulator.solver", "simulator.solver.matbuilder", "simulator.solver.PDESolver_IMP",
        "simulator.solver.PDESolver_EXP", "simulator.optimizer", "simulator.optimizer.tso",
        "analyzer", "analyzer.Data1D", "analyzer.Data1D.core1D", "analyzer.Data2D", "analyzer.Data2D.core2D",
        "utils", "utils.mesh_utils", "utils.history_utils",
    ]
    root = sys.modules.setdefault("fiberis", types.ModuleType("fiberis"))
    root.__path__ = []  # type: ignore[attr-defined]
    for n in names:
        mod = importlib.import_module(f"{__name__}.{n}")
        sys.modules[f"fiberis.{n}"] = mod
        parent, _, leaf = f"fiberis.{n}".rpartition(".")
        setattr(sys.modules[parent], leaf, mod)

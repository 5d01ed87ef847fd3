"""Import shim for the UNMODIFIED reference (fibeRIS) when it is mounted at /root/reference.

TEST INFRASTRUCTURE ONLY.  This module exists so that (a) ``oracle/gen_golden.py`` can run the real
reference in the build container to produce the committed fixtures under ``tests/golden/`` and
(b) the optional ``tests/test_oracle_vs_reference.py`` can compare the oracle with the live
reference when (and only when) ``/root/reference`` is present.  (c) ``bench.py --impl reference`` and ``tests/test_reference_suite.py`` can use the byte-identical copies
that ``oracle/make_ref.py`` places in the git-ignored ``oracle/_ref/`` (``/root/reference`` does not exist
on the GPU box).  Nothing in ``fiberis_b200/`` imports it.

The reference imports ``matplotlib`` at module top (``src/fiberis/simulator/core/pds.py:5``,
``src/fiberis/analyzer/Data1D/core1D.py:6``) and matplotlib is not installed in this image, so a
handful of empty stub modules are registered in ``sys.modules`` first.  No reference code is
copied; it is imported from where it lies.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = os.environ.get("FIBERIS_REFERENCE_SRC", "/root/reference/src")
# byte-identical copies made by oracle/make_ref.py (git-ignored; they travel to the GPU box)
TRAVEL_SRC = os.path.join(_HERE, "_ref", "src")


def _has(src) -> bool:
    return os.path.isdir(os.path.join(src, "fiberis", "simulator"))


def reference_src():
    """Where the unmodified reference can be imported from: the mounted tree, else the travelling copy."""
    if _has(REFERENCE_SRC):
        return REFERENCE_SRC
    if _has(TRAVEL_SRC):
        return TRAVEL_SRC
    return None


def reference_available() -> bool:
    return reference_src() is not None


class _Anything:
    """Attribute sink standing in for matplotlib classes used only in type hints / plotting."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):  # ``fig, ax = plt.subplots()``
        return iter((_Anything(), _Anything()))


def _install_matplotlib_stub() -> None:
    try:
        import matplotlib  # noqa: F401  (a real one wins)
        return
    except Exception:
        pass
    names = [
        "matplotlib", "matplotlib.pyplot", "matplotlib.axes", "matplotlib.lines", "matplotlib.dates",
        "matplotlib.ticker", "matplotlib.image", "matplotlib.collections", "matplotlib.figure",
        "matplotlib.colors", "matplotlib.cm", "matplotlib.patches", "matplotlib.gridspec",
        "matplotlib.animation", "matplotlib.colorbar", "matplotlib.widgets",
        "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1",
    ]
    for n in names:
        m = types.ModuleType(n)
        # classes (Axes, Line2D, ...) resolve to the sink class, functions (plt.subplots, ...) to a callable sink object
        m.__getattr__ = lambda attr, _n=n: _Anything if attr[:1].isupper() else _Anything()  # type: ignore[attr-defined]
        m.__path__ = []  # behave like a package so sub-imports resolve
        sys.modules.setdefault(n, m)
    sys.modules["matplotlib"].use = lambda *a, **k: None  # type: ignore[attr-defined]


def import_reference():
    """Return the reference's (pds, matbuilder, tso, Data1D, Data2D, PDESolver_IMP, PDESolver_EXP)."""
    src = reference_src()
    if src is None:
        raise ImportError(f"reference neither mounted at {REFERENCE_SRC} nor copied to {TRAVEL_SRC} (oracle/make_ref.py)")
    _install_matplotlib_stub()
    if src not in sys.path:
        sys.path.insert(0, src)
    from fiberis.simulator.core import pds  # type: ignore
    from fiberis.simulator.solver import matbuilder, PDESolver_IMP, PDESolver_EXP  # type: ignore
    from fiberis.simulator.optimizer import tso  # type: ignore
    from fiberis.analyzer.Data1D.core1D import Data1D  # type: ignore
    from fiberis.analyzer.Data2D.core2D import Data2D  # type: ignore
    return types.SimpleNamespace(pds=pds, matbuilder=matbuilder, tso=tso, Data1D=Data1D, Data2D=Data2D,
                                 PDESolver_IMP=PDESolver_IMP, PDESolver_EXP=PDESolver_EXP)

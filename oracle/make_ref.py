"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the simulator path, copied from where
they lie under ``/root/reference`` so that they can travel to the GPU box (which has no /root/reference).

TEST / BASELINE INFRASTRUCTURE — NOT PRODUCT CODE.  ``oracle/_ref/`` is git-ignored (no reference
source ever enters this repository's history) but not gpurun-ignored.  Two users:

* ``tests/test_reference_suite.py`` runs the reference's own three simulator test files
  (``tests/test_simulator_{core,solver,optimizer}.py``, byte-identical copies) against
  ``fiberis_b200.install_as_fiberis()`` — the drop-in proof;
* ``bench.py --impl reference`` times the reference's own ``PDS1D_*.solve`` (``kind: "reference"``).

The reference is pure Python; "building" it is copying the files its import graph needs for this path
(`fiberis.simulator.*`, the two result containers and the utils modules they import).  matplotlib is
not installed in this image: ``oracle/_refshim.py`` registers stub modules before the import.

    python oracle/make_ref.py            # (re)creates oracle/_ref/ when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("FIBERIS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# (path relative to the reference root) — everything is copied verbatim
FILES = [
    "src/fiberis/__init__.py",
    "src/fiberis/simulator/core/__init__.py",
    "src/fiberis/simulator/core/pds.py",
    "src/fiberis/simulator/core/bcs.py",
    "src/fiberis/simulator/solver/__init__.py",
    "src/fiberis/simulator/solver/matbuilder.py",
    "src/fiberis/simulator/solver/PDESolver_IMP.py",
    "src/fiberis/simulator/solver/PDESolver_EXP.py",
    "src/fiberis/simulator/optimizer/__init__.py",
    "src/fiberis/simulator/optimizer/tso.py",
    "src/fiberis/analyzer/__init__.py",
    "src/fiberis/analyzer/Data1D/__init__.py",
    "src/fiberis/analyzer/Data1D/core1D.py",
    "src/fiberis/analyzer/Data2D/__init__.py",
    "src/fiberis/analyzer/Data2D/core2D.py",
    "src/fiberis/utils/__init__.py",
    "src/fiberis/utils/history_utils.py",
    "src/fiberis/utils/signal_utils.py",
    "src/fiberis/utils/mesh_utils.py",
    "tests/test_simulator_core.py",
    "tests/test_simulator_solver.py",
    "tests/test_simulator_optimizer.py",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "fiberis", "simulator"))


def make(verbose: bool = False) -> bool:
    """Copy the files (only when the reference is mounted).  Returns True when oracle/_ref/ is usable."""
    if not available():
        return os.path.isfile(os.path.join(OUT, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(src):  # a package marker the reference does not have (namespace package)
            if rel.endswith("__init__.py"):
                continue
            raise FileNotFoundError(src)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
        if verbose:
            print("copied", rel)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_ROOT, "files": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = make(verbose=True)
    print("oracle/_ref ready" if ok else "reference not mounted and no previous copy")
    sys.exit(0 if ok else 1)

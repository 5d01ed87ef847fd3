"""The reference's OWN simulator tests, unmodified, as the acceptance suite.

``oracle/make_ref.py`` copies ``/root/reference/tests/test_simulator_{core,solver,optimizer}.py`` byte for byte
into the git-ignored ``oracle/_ref/tests/`` (they travel to the GPU box).  Here they run in a subprocess

* against the reference itself (CPU; checks the copy and the matplotlib stub), and
* against ``fiberis_b200.install_as_fiberis()`` on the GPU: every ``import fiberis.simulator...`` in those files
  resolves to this package, whose solves run through the C ABI on the device.
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")
FILES = ["test_simulator_core.py", "test_simulator_solver.py", "test_simulator_optimizer.py"]
N_TESTS = 59  # what the three files hold (SURVEY.md 8(c) probe)


def _have_copy():
    return all(os.path.isfile(os.path.join(REF_TESTS, f)) for f in FILES)


def _run(plugin):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", plugin, "-p", "no:cacheprovider", "--rootdir", REF_TESTS,
           "-c", os.devnull] + [os.path.join(REF_TESTS, f) for f in FILES]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    m = re.search(r"(\d+) passed", r.stdout)
    return r.returncode, int(m.group(1)) if m else 0, tail


@pytest.mark.skipif(not _have_copy(), reason="oracle/_ref not built (python oracle/make_ref.py needs /root/reference)")
def test_copy_passes_against_the_reference_itself():
    rc, passed, tail = _run("tests._refself_plugin")
    assert rc == 0 and passed == N_TESTS, tail


@pytest.mark.gpu
def test_unmodified_reference_tests_pass_on_the_b200_package():
    assert _have_copy(), "oracle/_ref/tests missing: run python oracle/make_ref.py where /root/reference is mounted"
    rc, passed, tail = _run("tests._dropin_plugin")
    assert rc == 0 and passed == N_TESTS, tail

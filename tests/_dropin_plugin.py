"""pytest plugin (``-p tests._dropin_plugin``): makes ``import fiberis.simulator...`` resolve to fiberis_b200 BEFORE
the reference's unmodified test files are collected — the drop-in proof of tests/test_reference_suite.py."""
import fiberis_b200

fiberis_b200.install_as_fiberis()

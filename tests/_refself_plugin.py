"""pytest plugin (``-p tests._refself_plugin``): puts the unmodified reference (mounted tree or the oracle/_ref copy) on
sys.path with the matplotlib stub, so that its own test files run against ITSELF (sanity of the copy)."""
from oracle import _refshim

_refshim.import_reference()

"""Import shim: the package directory is named after the reference repository
(`mixed-integer-optimal-control---algorithm-tools_b200`), which is not a valid Python identifier, so it is
loaded through importlib and re-exported here as `mioc_b200`."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("mixed-integer-optimal-control---algorithm-tools_b200")
sys.modules[__name__] = _pkg

"""Import alias: `import sos_b200` -> the package in ./sos-radiative-transfer_b200/.

The package directory carries the name the project brief prescribes, which contains hyphens
and therefore cannot appear in an `import` statement.
"""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("sos-radiative-transfer_b200")
sys.modules[__name__] = _pkg

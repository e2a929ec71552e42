"""Import alias: the package directory is named ``ground-plane-polling_b200`` (not a Python identifier), so
``import gpp_b200`` resolves to it.  Use ``from gpp_b200 import fit_road_planes`` or attribute access
(``gpp_b200.layers.fit_road_planes``); do not ``import gpp_b200.layers`` as a dotted module path."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module('ground-plane-polling_b200')
sys.modules[__name__] = _pkg

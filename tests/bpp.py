"""Imports the product package (its directory name contains a hyphen, so importlib is needed)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
pkg = importlib.import_module("bulletproofs-plus_b200")
ffi = importlib.import_module("bulletproofs-plus_b200._ffi")

_engine = None


def engine():
    """One Engine per test process (GPU tests only)."""
    global _engine
    if _engine is None:
        _engine = pkg.Engine(0)
    return _engine

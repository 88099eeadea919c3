"""Importable alias of the `learning-based-rgb-d-image-compression_b200` package."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
sys.modules[__name__] = importlib.import_module("learning-based-rgb-d-image-compression_b200")

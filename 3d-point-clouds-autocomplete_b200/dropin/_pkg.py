"""Locates the implementation package for the drop-in modules."""
import importlib
import os
import sys

_repo_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _repo_root not in sys.path:
    sys.path.insert(0, _repo_root)
pkg = importlib.import_module("3d-point-clouds-autocomplete_b200")

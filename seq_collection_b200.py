"""Importable alias for the `seq-collection_b200/` package (its directory name has a hyphen)."""
import importlib
import sys

_pkg = importlib.import_module("seq-collection_b200")
sys.modules[__name__] = _pkg

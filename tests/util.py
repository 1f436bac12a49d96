"""Shared helpers for the tests: engine import (the package directory name has hyphens) and workloads."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def engine_pkg():
    return importlib.import_module("sdr-j-dab_b200")

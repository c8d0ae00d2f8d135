"""Helpers for the GPU parity tests: load the product package (hyphenated directory name)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pkg = importlib.import_module("kimera-rpgo_b200")
synth = importlib.import_module("kimera-rpgo_b200.synth")
PcmGpu = pkg.PcmGpu

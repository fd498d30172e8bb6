"""
Imports the unmodified reference from baseline/_ref (see install_reference.py). TEST / BENCHMARK INFRASTRUCTURE: used by
`bench.py --impl reference` (the CPU arm) and by the drop-in tests; nothing under shot_fpfh_b200/ imports this.

The reference's package `__init__` pulls in matplotlib (analysis/*.py, pca_based_descriptors.py:9) and its CLI pulls in
coloredlogs (register_point_clouds.py:8); neither is installed in this image and neither is used on the hot path, so
empty stub modules stand in for them.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "shot_fpfh"))


def load():
    """-> the reference's top-level `shot_fpfh` module. Raises RuntimeError when baseline/_ref is absent."""
    if not available():
        raise RuntimeError(f"{REF_DIR} is missing: run `python baseline/install_reference.py` in the build container")
    for name in ("matplotlib", "matplotlib.pyplot", "coloredlogs"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    plt = sys.modules["matplotlib.pyplot"]
    for fn in ("figure", "show", "subplots", "savefig", "close", "plot", "hist", "legend", "title"):
        if not hasattr(plt, fn):
            setattr(plt, fn, lambda *a, **k: None)
    cl = sys.modules["coloredlogs"]
    if not hasattr(cl, "install"):
        cl.install = lambda *a, **k: None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return importlib.import_module("shot_fpfh")

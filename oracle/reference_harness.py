"""
TEST INFRASTRUCTURE ONLY. Imports the unmodified reference from /root/reference.

Only usable in the build container: /root/reference does not exist on the GPU box, so nothing that runs
there (`-m gpu` tests, smoke(), bench.py) may call this. It is used by `oracle/make_golden.py` and by the
CPU tests that are skipped when the reference is absent.

The reference's package `__init__` pulls in matplotlib (analysis/*.py, pca_based_descriptors.py:9) and the
CLI pulls in coloredlogs (register_point_clouds.py:8); neither is installed here and neither is used on the
hot path, so empty stub modules are enough.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "shot_fpfh"))


def _install_stubs() -> None:
    for name in ("matplotlib", "matplotlib.pyplot", "coloredlogs"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]


def import_reference():
    """Returns the reference's top-level `shot_fpfh` module (unmodified sources, read in place)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("shot_fpfh")


def patched_fpfh_decorrelated():
    """
    Returns a `compute_fpfh_descriptor` whose `decorrelated=True` branch runs: the reference assigns a
    (n_bins, 3) array into a (3 * n_bins,) row (fpfh.py:59-79) and raises ValueError. The patch is the
    single token `).T` -> `).ravel()` at fpfh.py:78, applied to the source text in memory (nothing is
    written anywhere). Layout of the result: [alpha bins | phi bins | theta bins].
    """
    import_reference()
    path = os.path.join(REFERENCE_ROOT, "shot_fpfh", "descriptors", "fpfh.py")
    with open(path) as f:
        lines = f.read().split("\n")
    assert lines[77].strip() == ").T", f"unexpected reference text at fpfh.py:78: {lines[77]!r}"
    lines[77] = lines[77].replace(").T", ").ravel()")
    module = types.ModuleType("shot_fpfh_fpfh_patched")
    exec(compile("\n".join(lines), path + " (patched :78)", "exec"), module.__dict__)
    return module.compute_fpfh_descriptor

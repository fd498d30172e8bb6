"""Mirror of `shot_fpfh.descriptors` for the hot path: the names pipeline.py imports (pipeline.py:15)."""

from .fpfh import compute_fpfh_descriptor
from .shot_parallelization import ShotMultiprocessor

__all__ = ["compute_fpfh_descriptor", "ShotMultiprocessor"]

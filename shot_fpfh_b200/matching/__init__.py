"""Mirror of `shot_fpfh.matching` for the hot path: the names pipeline.py imports (pipeline.py:24-30), minus RANSAC."""

from .filters import FilterFunction, left_median_filter, quantile_filter, threshold_filter
from .matching import basic_matching, double_matching_with_rejects, match_descriptors

__all__ = [
    "FilterFunction",
    "threshold_filter",
    "quantile_filter",
    "left_median_filter",
    "match_descriptors",
    "basic_matching",
    "double_matching_with_rejects",
]

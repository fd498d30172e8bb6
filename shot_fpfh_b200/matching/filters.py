"""
Filters on the nearest-neighbour distance vector (reference: shot_fpfh/matching/filters.py:12-40). They are O(Q)
host-side masks applied to the float64 distances the re-rank kernel returns, and stay NumPy.
"""

from __future__ import annotations

from typing import Any, Protocol

import numpy as np
import numpy.typing as npt


class FilterFunction(Protocol):
    def __call__(self, distances: npt.NDArray[np.float64], *args: Any, **kwargs: Any) -> npt.NDArray[np.bool_]: ...


def threshold_filter(distances: npt.NDArray[np.float64], threshold_multiplier: float) -> npt.NDArray[np.bool_]:
    """Keeps the matches closer than `threshold_multiplier` times the smallest non-zero distance (filters.py:19-23)."""
    smallest = distances[distances != 0].min()
    return distances <= smallest * threshold_multiplier


def quantile_filter(distances: npt.NDArray[np.float64], quantiles: tuple[float, float]) -> npt.NDArray[np.bool_]:
    """Keeps the matches whose distance lies between two quantiles (filters.py:26-31)."""
    low, high = np.quantile(distances, quantiles)
    return (distances >= low) & (distances <= high)


def left_median_filter(distances: npt.NDArray[np.float64]) -> npt.NDArray[np.bool_]:
    """
    filters.py:34-40, kept literally: the lower bound mixes the median distance with the smallest INDEX of a
    non-zero distance (`distances.nonzero()[0].min()`), which is what the reference computes.
    """
    median = np.median(distances)
    return (distances <= median) & (distances >= (median + distances.nonzero()[0].min()) / 2)

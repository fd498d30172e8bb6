"""
TEST INFRASTRUCTURE — CPU restatement of the reference's voxel-based keypoint selectors.
Only tests/ may import this module; the product never does.

  select_keypoints_subsampling             shot_fpfh/keypoint_selection.py:34-46  (= core/subsampling.py:5-39)
  select_keypoints_with_density_threshold  shot_fpfh/keypoint_selection.py:65-122

Pinned by tests/test_oracle_golden.py::test_keypoint_oracle_against_reference, which runs the unmodified reference
functions on seeded clouds in the build container and compares index for index.
"""

from __future__ import annotations

import numpy as np
from sklearn.neighbors import KDTree


def _voxels(points: np.ndarray, voxel_size: float):
    """
    Occupied voxels in lexicographic key order -> list of member index arrays, in the order the reference visits
    them: `np.argsort(inverse)` with NumPy's default (unstable) sort, keypoint_selection.py:85 — the order decides
    the rounding of the barycentre and which of two equidistant members is "first".
    """
    keys = ((points - points.min(axis=0)) // voxel_size).astype(int)  # keypoint_selection.py:78 / subsampling.py:13
    _, inverse, counts = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
    order = np.argsort(inverse.ravel())
    bounds = np.concatenate([[0], np.cumsum(counts)])
    return [order[bounds[v] : bounds[v + 1]] for v in range(counts.shape[0])]


def _representative(points: np.ndarray, members: np.ndarray) -> int:
    """The member closest to the voxel's barycentre, first one on ties (keypoint_selection.py:98-103)."""
    centre = points[members].mean(axis=0)
    return int(members[np.linalg.norm(points[members] - centre, axis=1).argmin()])


def select_keypoints_subsampling(points: np.ndarray, voxel_size: float) -> np.ndarray:
    return np.array([_representative(points, m) for m in _voxels(points, voxel_size)], dtype=np.int64)


def select_keypoints_with_density_threshold(points, voxel_size, density_threshold_value, density_threshold_radius=None):
    if density_threshold_radius is None:
        density_threshold_radius = voxel_size  # :82-83
    tree = KDTree(points) if density_threshold_radius != voxel_size else None  # :88-89
    kept = []
    for members in _voxels(points, voxel_size):
        rep = _representative(points, members)
        if tree is None:
            dense = members.shape[0] > density_threshold_value  # :105-106
        else:  # :108-115 — the count includes the representative itself
            dense = tree.query_radius([points[rep]], density_threshold_radius)[0].shape[0] > density_threshold_value
        if dense:
            kept.append(rep)
    return np.array(kept)

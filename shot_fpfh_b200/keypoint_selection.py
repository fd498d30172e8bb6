"""
Keypoint selection with the reference's signatures (shot_fpfh/keypoint_selection.py), for the two selectors that
are voxel-grid reductions of the cloud: they sit immediately upstream of the descriptors and
are Python loops over voxels in the reference (4.4 s at 1M points).

  select_keypoints_subsampling(points, voxel_size)                       keypoint_selection.py:34-46
  select_keypoints_with_density_threshold(points, voxel_size, value, radius=None)        :65-122

Both run on the device (csrc/subsample.cu; the density variant with a radius different from the voxel size also
uses the uniform grid's fixed-radius count, csrc/grid.cu). The iterative and the random selectors of the reference
are not data-parallel (a sequential greedy cover; NumPy's global RNG) and are left to the reference.
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt


def select_keypoints_subsampling(points: npt.NDArray[np.float64], voxel_size: float) -> npt.NDArray[np.int64]:
    """Indices of the points closest to the barycentre of each occupied voxel (= `grid_subsampling`)."""
    from .subsampling import grid_subsampling_gpu

    return grid_subsampling_gpu(points, voxel_size)


def select_keypoints_with_density_threshold(
    points: npt.NDArray[np.float64],
    voxel_size: float,
    density_threshold_value: int,
    density_threshold_radius: float | None = None,
) -> npt.NDArray[np.int64]:
    """
    The voxel representatives of `select_keypoints_subsampling`, kept only where the density exceeds the threshold:
    more than `density_threshold_value` points in the voxel when the radius is the voxel size (or None), else more
    than that many points of the cloud within `density_threshold_radius` of the representative (itself included,
    as `KDTree.query_radius` counts it) — keypoint_selection.py:104-118. Lexicographic voxel order, like the
    reference's `np.unique(axis=0)`.
    """
    import torch

    from . import ops
    from .device import Grid, upload

    pts = upload(points)
    picked, members = ops.voxel_subsample(pts, float(voxel_size), want_members=True)
    if picked.shape[0] == 0:
        return np.array([])  # what `np.array([])` of the reference's empty list is
    if density_threshold_radius is None or density_threshold_radius == voxel_size:
        keep = members > int(density_threshold_value)
    else:
        grid = Grid().build(pts, None, float(density_threshold_radius))
        offsets, _, _, _ = ops.radius_csr(grid, pts[picked].contiguous(), float(density_threshold_radius), count_only=True)
        keep = (offsets[1:] - offsets[:-1]) > int(density_threshold_value)
        torch.cuda.synchronize()
        grid.close()
    out = picked[keep].cpu().numpy()
    return out if out.shape[0] else np.array([])

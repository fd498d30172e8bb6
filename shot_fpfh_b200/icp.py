"""
`icp_point_to_plane` with the reference's signature (shot_fpfh/icp.py:137-189).

Per iteration the reference moves the subsampled scan, asks a KD-tree for each point's nearest reference point,
drops the pairs farther than `d_max`, builds the point-to-plane system `g^T g x = g^T h` (core/solvers.py:34-48),
composes the step and reports the mean residual. Here one device call per iteration does everything up to the
29 sums of that system (`sf_icp_plane_step`, csrc/registration.cu: warp per point on a uniform grid over the
reference cloud, fixed summation order); the 6x6 solve, the Euler step and the composition are the reference's
host arithmetic on those sums.

`icp_point_to_point` (icp.py:81-134) is provided as what its code and docstring intend: the reference's own line for
the RMS (icp.py:118-120) subtracts `ref[neighbors]` of shape (n, 1, 3) from the (n, 3) inliers — an (n, n, 3) broadcast
that leaves an array of three numbers, which the next line cannot format: TypeError on every input with more than one
pair (observed with the unmodified reference). The intended
quantity is the one its sibling `icp_point_to_point_with_sampling` computes (icp.py:68-72): the root of the SUM of the
squared distances between the inliers and their nearest reference points. The nearest-neighbour search of every
iteration (`KDTree.query`) runs on the device; the 3x3 Kabsch solve and the composition are the reference's host
arithmetic. `icp_point_to_point_with_sampling` draws from NumPy's global unseeded generator and stays the reference's.
"""

from __future__ import annotations

import logging

import numpy as np
import numpy.typing as npt

from .core import RigidTransform
from .core.solvers import solver_point_to_point, transform_from_plane_system


def icp_point_to_point(
    scan: npt.NDArray[np.float64],
    ref: npt.NDArray[np.float64],
    transformation_init: RigidTransform,
    d_max: float,
    voxel_size: float = 0.2,
    max_iter: int = 100,
    rms_threshold: float = 1e-2,
    disable_progress_bar: bool = False,
) -> tuple[RigidTransform, float, bool]:
    """Returns (transformation, root of the summed squared inlier distances of the last iteration, whether it fell
    below `rms_threshold`). Signature of icp.py:81-90."""
    import torch

    from . import ops
    from .device import Grid, upload

    ref = np.ascontiguousarray(ref, dtype=np.float64)
    ref_dev, scan_dev = upload(ref), upload(scan)
    # (the search kernel is the point-to-plane step's: it wants normals on the grid; its plane sums are not used here)
    grid = Grid().build(ref_dev, torch.zeros_like(ref_dev), float(d_max))
    subsampled_dev = scan_dev[ops.voxel_subsample(scan_dev, float(voxel_size))].contiguous()  # icp.py:95
    subsampled = subsampled_dev.cpu().numpy()
    transformation_icp = RigidTransform(np.asarray(transformation_init.rotation, dtype=np.float64),
                                        np.asarray(transformation_init.translation, dtype=np.float64))
    rms = 0.0
    try:
        for _ in range(max_iter):
            _, nearest = ops.icp_plane_step(grid, subsampled_dev, transformation_icp.as_row(), float(d_max), want_nearest=True)
            nearest = nearest.cpu().numpy()
            keep = nearest >= 0  # pairs within d_max (icp.py:109-110)
            inliers = transformation_icp[subsampled[keep]]
            targets = ref[nearest[keep]]
            step = solver_point_to_point(inliers, targets)  # raises without inliers, like the reference
            rms = float(np.sqrt((np.linalg.norm(inliers - targets, axis=1) ** 2).sum(axis=0)))
            transformation_icp = step @ transformation_icp
            if rms < rms_threshold:
                logging.info("RMS threshold reached.")
                break
    finally:
        grid.close()
    return transformation_icp, rms, rms < rms_threshold


def icp_point_to_plane(
    scan: npt.NDArray[np.float64],
    ref: npt.NDArray[np.float64],
    ref_normals: npt.NDArray[np.float64],
    transformation_init: RigidTransform,
    d_max: float,
    voxel_size: float = 0.2,
    max_iter: int = 50,
    rms_threshold: float = 1e-2,
    disable_progress_bar: bool = False,
) -> tuple[RigidTransform, float, bool]:
    """Returns (transformation, last mean point-to-plane residual, whether it fell below `rms_threshold`)."""
    from . import ops
    from .device import Grid, upload

    ref_dev, normals_dev, scan_dev = upload(ref), upload(ref_normals), upload(scan)
    grid = Grid().build(ref_dev, normals_dev, float(d_max))
    subsampled = scan_dev[ops.voxel_subsample(scan_dev, float(voxel_size))].contiguous()  # icp.py:156
    # any object with .rotation / .translation (the reference's own RigidTransform under dropin, the ground truth of
    # get_transform_from_conf_file, ...) is taken as the initial transformation
    transformation_icp = RigidTransform(np.asarray(transformation_init.rotation, dtype=np.float64),
                                        np.asarray(transformation_init.translation, dtype=np.float64))
    rms = 0.0
    upper = np.triu_indices(6)
    try:
        for _ in range(max_iter):
            sums = ops.icp_plane_step(grid, subsampled, transformation_icp.as_row(), float(d_max))
            gtg = np.zeros((6, 6))
            gtg[upper] = sums[:21]
            gtg = gtg + np.triu(gtg, 1).T
            step = transform_from_plane_system(gtg, sums[21:27])  # raises LinAlgError without inliers, like the reference
            transformation_icp = step @ transformation_icp
            rms = sums[27] / sums[28]
            if rms < rms_threshold:
                logging.info("RMS threshold reached.")
                break
    finally:
        grid.close()
    return transformation_icp, rms, rms < rms_threshold

"""
The two closed-form solvers of the reference (shot_fpfh/core/solvers.py:9-48), host NumPy: they act on a handful of
points (RANSAC draws) or on sums the device has already reduced (ICP), so there is nothing to move to the GPU.
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt
from scipy.spatial.transform import Rotation

from .rigid_transform import RigidTransform


def solver_point_to_point(scan: npt.NDArray[np.float64], ref: npt.NDArray[np.float64]) -> RigidTransform:
    """Least-squares rigid transform mapping `scan` onto `ref` (Kabsch by SVD, solvers.py:9-31)."""
    scan_centre, ref_centre = scan.mean(axis=0), ref.mean(axis=0)
    u, _, vt = np.linalg.svd((scan - scan_centre).T.dot(ref - ref_centre))
    rotation = vt.T @ u.T
    if np.linalg.det(rotation) < 0:  # a reflection: flip the direction of the smallest singular value
        ut = u.T
        ut[-1] *= -1
        rotation = vt.T @ ut
    return RigidTransform(rotation, ref_centre - rotation.dot(scan_centre))


def transform_from_plane_system(gtg: npt.NDArray[np.float64], gth: npt.NDArray[np.float64]) -> RigidTransform:
    """The step of solvers.py:45-48 from the 6x6 normal equations: small-angle rotation (xyz Euler) + translation."""
    solution = np.linalg.solve(gtg, gth)
    return RigidTransform(Rotation.from_euler("xyz", solution[:3]).as_matrix(), solution[3:6])


def solver_point_to_plane(scan, ref, normals_ref) -> RigidTransform:
    """Linearised point-to-plane step (solvers.py:34-48)."""
    g = np.hstack((np.cross(scan, normals_ref), normals_ref))
    h = np.einsum("ij, ij->i", ref - scan, normals_ref)
    return transform_from_plane_system(g.T @ g, g.T @ h)

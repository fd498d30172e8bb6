"""
TEST INFRASTRUCTURE — CPU restatement of the reference's RANSAC and ICP.
Only tests/ may import this module; the product never does.

  ransac_on_matches    shot_fpfh/matching/ransac.py:17-82   (generator passed in: the reference keeps a module-level
                       `default_rng(72)`, so a fresh `default_rng(72)` here reproduces its FIRST call)
  icp_point_to_plane   shot_fpfh/icp.py:137-189
  icp_point_to_point   shot_fpfh/icp.py:81-134 as intended (the reference raises: see the function)
  solvers              shot_fpfh/core/solvers.py:9-48, rigid transform algebra core/rigid_transform.py:45-70

Pinned by tests/test_oracle_golden.py::test_registration_oracle_against_reference.
Transforms are returned as (rotation (3,3), translation (3,)).
"""

from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation
from sklearn.neighbors import KDTree


def kabsch(scan, ref):  # solvers.py:9-31
    cs, cr = scan.mean(axis=0), ref.mean(axis=0)
    u, _, vt = np.linalg.svd((scan - cs).T.dot(ref - cr))
    rot = vt.T @ u.T
    if np.linalg.det(rot) < 0:
        ut = u.T
        ut[-1] *= -1
        rot = vt.T @ ut
    return rot, cr - rot.dot(cs)


def normalized(rot):  # rigid_transform.py:45-52
    q = Rotation.from_matrix(rot).as_quat()
    return Rotation.from_quat(q / np.linalg.norm(q)).as_matrix()


def compose(a, b):  # a after b, rigid_transform.py:54-70
    return normalized(a[0] @ b[0]), a[0] @ b[1] + a[1]


def apply(t, points):  # rigid_transform.py:81-88
    return points.dot(t[0].T) + t[1]


def ransac_on_matches(scan_idx, ref_idx, scan_kp, ref_kp, rng, n_draws=10000, draw_size=4, distance_threshold=1.0):
    best_n, best_t = None, None
    for _ in range(n_draws):
        draw = rng.choice(scan_idx.shape[0], draw_size, replace=False, shuffle=False)  # :48-53
        t = kabsch(scan_kp[scan_idx[draw]], ref_kp[ref_idx[draw]])
        n_in = (np.linalg.norm(apply(t, scan_kp[scan_idx]) - ref_kp[ref_idx], axis=1) <= distance_threshold).sum()  # :55-62
        if best_n is None or n_in > best_n:  # :63
            best_n, best_t = n_in, t
    return best_n / scan_idx.shape[0], (normalized(best_t[0]), best_t[1])  # :78-80


def point_to_plane_step(scan, ref, normals):  # solvers.py:34-48
    g = np.hstack((np.cross(scan, normals), normals))
    h = np.einsum("ij, ij->i", ref - scan, normals)
    sol = np.linalg.solve(g.T @ g, g.T @ h)
    return Rotation.from_euler("xyz", sol[:3]).as_matrix(), sol[3:6]


def icp_point_to_point(scan, ref, init, d_max, subsampled_indices, max_iter=100, rms_threshold=1e-2):
    """icp.py:81-134 as intended: its RMS line (:118-120) mixes shapes (n, 3) and (n, 1, 3) and the reference raises a
    TypeError on every input with more than one pair, so this restatement cannot be pinned to reference outputs; the
    quantity is the one of icp.py:68-72, the root of the summed squared inlier distances."""
    tree = KDTree(ref)
    t = init
    rms = 0.0
    iterations = 0
    for _ in range(max_iter):
        iterations += 1
        aligned = apply(t, scan[subsampled_indices])  # :103
        dist, nn = tree.query(aligned)  # :104
        keep = dist.squeeze(axis=1) <= d_max
        inliers, nbrs = aligned[keep], nn[keep, 0]
        step = kabsch(inliers, ref[nbrs])  # :113-115
        rms = np.sqrt((np.linalg.norm(inliers - ref[nbrs], axis=1) ** 2).sum(axis=0))
        t = compose(step, t)  # :121
        if rms < rms_threshold:
            break
    return t, rms, rms < rms_threshold, iterations


def icp_point_to_plane(scan, ref, ref_normals, init, d_max, subsampled_indices, max_iter=50, rms_threshold=1e-2):
    """`subsampled_indices`: what `grid_subsampling(scan, voxel_size)` returns at icp.py:156 (passed in so that the
    test can hand both sides the same selection; its ties are a separate, documented matter)."""
    tree = KDTree(ref)
    t = init
    rms = 0.0
    iterations = 0
    for _ in range(max_iter):
        iterations += 1
        aligned = apply(t, scan[subsampled_indices])
        dist, nn = tree.query(aligned)  # :158
        keep = dist.squeeze(axis=1) <= d_max
        inliers, nbrs = aligned[keep], nn[keep, 0]
        step = point_to_plane_step(inliers, ref[nbrs], ref_normals[nbrs])
        t = compose(step, t)  # :174
        rms = np.abs(np.einsum("ij, ij->i", inliers - ref[nbrs], ref_normals[nbrs])).mean(axis=0)  # :175-182
        if rms < rms_threshold:
            break
    return t, rms, rms < rms_threshold, iterations

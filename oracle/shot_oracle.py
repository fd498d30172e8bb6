"""
TEST INFRASTRUCTURE ONLY. NumPy restatement of the reference's SHOT path.

Follows shot_fpfh/descriptors/shot.py:16-48 (local reference frame), :51-70 (azimuth octant), :73-118
(radial interpolation), :121-171 (elevation interpolation), :175-306 (descriptor) and the single-scale
driver shot_fpfh/descriptors/shot_parallelization.py:135-183.

The reference's descriptor is NOT an accumulating histogram: its ten scatter statements are
NumPy fancy-index `descriptor[idx] += values`, which gathers, adds and scatters, so that when several
neighbours address the same bin in one statement only the LAST one (largest distance, the neighbours being
sorted by ascending distance at shot.py:218-221) contributes — even when its value is 0. This file states
that rule explicitly ("winner selection") instead of relying on the buffering side effect, with the same
float64 arithmetic, so that it reproduces the reference bit for bit (checked by oracle/make_golden.py and
tests/test_oracle_golden.py).
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt
from sklearn.neighbors import KDTree

N_COS, N_AZ, N_EL, N_RAD = 11, 8, 2, 2
SHOT_LEN = N_COS * N_AZ * N_EL * N_RAD  # 352


def local_reference_frame(
    point: npt.NDArray[np.float64], neighbors: npt.NDArray[np.float64], radius: float
) -> npt.NDArray[np.float64]:
    """
    shot.py:23-48. Weighted covariance with weights (radius - distance) over ALL the neighbours handed in
    (the multiprocess driver includes the query point itself), eigenvectors by ascending
    eigenvalue, x = largest / z = smallest, each flipped when strictly more neighbours project negatively
    than non-negatively, y = z cross x. Columns of the result are [x, y, z].
    """
    if neighbors.shape[0] == 0:
        return np.eye(3)
    rel = neighbors - point
    weights = radius - np.linalg.norm(rel, axis=1)
    cov = rel.T @ (rel * weights[:, None]) / weights.sum()
    _, vec = np.linalg.eigh(cov)
    x_axis, z_axis = vec[:, 2].copy(), vec[:, 0].copy()
    proj = rel @ x_axis
    if (proj < 0).sum() > (proj >= 0).sum():
        x_axis = -x_axis
    proj = rel @ z_axis
    if (proj < 0).sum() > (proj >= 0).sum():
        z_axis = -z_axis
    y_axis = np.cross(z_axis, x_axis)
    return np.stack((x_axis, y_axis, z_axis), axis=1)


def azimuth_octant(x: npt.NDArray[np.float64], y: npt.NDArray[np.float64]) -> npt.NDArray[np.int64]:
    """shot.py:60-70. Octant 0 starts at angle -pi; decided by comparisons only."""
    upper = (y > 0) | ((y == 0) & (x < 0))
    right = (x > 0) | ((x == 0) & (y > 0))
    second_half = np.where((x * y > 0) | (x == 0), np.abs(x) < np.abs(y), np.abs(x) > np.abs(y))
    return 4 * upper.astype(np.int64) + 2 * np.logical_xor(right, upper) + second_half


def radial_weights(rho: npt.NDArray[np.float64], radius: float):
    """shot.py:95-118 -> (towards outer shell, towards inner shell, own shell)."""
    half = radius / 2
    to_inner = ((rho > radius / 2) & (rho < radius * 3 / 4)) * (radius * 3 / 4 - rho) / half
    to_outer = ((rho < radius / 2) & (rho > radius / 4)) * (rho - radius / 4) / half
    own = (rho < radius / 2) * (1 - np.abs(rho - radius / 4) / half) + (rho > radius / 2) * (
        1 - np.abs(rho - radius * 3 / 4) / half
    )
    return to_outer, to_inner, own


def elevation_weights(phi: npt.NDArray[np.float64], z: npt.NDArray[np.float64]):
    """shot.py:142-171 -> (towards elevation bin 1, towards elevation bin 0, own bin)."""
    half_pi = np.pi / 2
    at_equator = np.abs(phi - np.pi / 2) < 1e-10
    to_upper = (((phi > np.pi / 2) | (at_equator & (z <= 0))) & (phi <= np.pi * 3 / 4)) * (
        np.pi * 3 / 4 - phi
    ) / half_pi
    to_lower = (((phi < np.pi / 2) & (~at_equator | (z > 0))) & (phi >= np.pi / 4)) * (
        phi - np.pi / 4
    ) / half_pi
    own = (phi < np.pi / 2) * (1 - np.abs(phi - np.pi / 4) / half_pi) + (phi >= np.pi / 2) * (
        1 - np.abs(phi - np.pi * 3 / 4) / half_pi
    )
    return to_upper, to_lower, own


def _flat(ci, ti, ei, ri):
    """C-order ravel of [cos 11][azimuth 8][elevation 2][radial 2] (shot.py:199-201, :303)."""
    return ((ci * N_AZ + ti) * N_EL + ei) * N_RAD + ri


def shot_writes(local: npt.NDArray[np.float64], cosine, rho, radius: float):
    """
    The ten (target bin, value) streams of shot.py:244-298 for neighbours already sorted by ascending rho.
    Returns a list of ten (int64[K], float64[K]) pairs in statement order.
    """
    x, y, z = local[:, 0], local[:, 1], local[:, 2]
    theta = np.arctan2(y, x)
    phi = np.arccos(np.clip(z / rho, -1, 1))

    cos_pos = (cosine + 1.0) * N_COS / 2.0 - 0.5
    ci = np.rint(cos_pos).astype(int)
    ti = azimuth_octant(x, y)
    ei = (z > 0).astype(int)
    ri = (rho > radius / 2).astype(int)

    d_cos = cos_pos - ci
    s_cos = np.sign(d_cos)
    a_cos = s_cos * d_cos
    own = _flat(ci, ti, ei, ri)
    writes = [
        (_flat((ci + s_cos).astype(int) % N_COS, ti, ei, ri), a_cos * ((ci > -0.5) & (ci < N_COS - 0.5))),
        (own, 1 - a_cos),
    ]

    to_outer, to_inner, own_shell = radial_weights(rho, radius)
    writes += [
        (_flat(ci, ti, ei, 1), to_outer * (ri == 0)),
        (_flat(ci, ti, ei, 0), to_inner * (ri == 1)),
        (own, own_shell),
    ]

    to_upper, to_lower, own_vol = elevation_weights(phi, z)
    writes += [
        (_flat(ci, ti, 1, ri), to_upper * (ei == 0)),
        (_flat(ci, ti, 0, ri), to_lower * (ei == 1)),
        (own, own_vol),
    ]

    az_size = 2 * np.pi / N_AZ
    d_az = np.clip((theta - (-np.pi + ti * az_size)) / az_size - 0.5, -0.5, 0.5)
    s_az = np.sign(d_az)
    a_az = s_az * d_az
    writes += [
        (_flat(ci, (ti + s_az).astype(int) % N_AZ, ei, ri), a_az),
        (own, 1 - a_az),
    ]
    return writes


def apply_last_writer(writes, n_neighbors: int) -> npt.NDArray[np.float64]:
    """
    desc[b] = sum over the ten statements of the value of the LAST neighbour (largest rank) that addressed
    bin b in that statement; 0 when nobody did. Statement order is kept so that the float64 additions
    happen in the reference's order.
    """
    desc = np.zeros(SHOT_LEN)
    rank = np.arange(n_neighbors)
    for target, value in writes:
        winner = np.full(SHOT_LEN, -1, dtype=np.int64)
        np.maximum.at(winner, target, rank)
        hit = winner >= 0
        desc[hit] = desc[hit] + value[winner[hit]]
    return desc


def apply_accumulate(writes) -> npt.NDArray[np.float64]:
    """The textbook histogram (every write accumulates). NOT what the reference computes (F1); kept to
    show in the tests that it is far from the reference."""
    desc = np.zeros(SHOT_LEN)
    for target, value in writes:
        np.add.at(desc, target, value)
    return desc


def shot_descriptor(
    point, neighbors, normals, radius: float, lrf, normalize: bool, min_neighborhood_size: int
) -> npt.NDArray[np.float64]:
    """shot.py:175-306 for one query (the per-task function of the multiprocess driver)."""
    rho = np.linalg.norm(neighbors - point, axis=1)
    keep = rho > 0
    if keep.sum() <= min_neighborhood_size:
        return np.zeros(SHOT_LEN)
    local = (neighbors[keep] - point) @ lrf
    cosine = np.clip(normals[keep] @ lrf[:, 2].T, -1, 1)
    rho = rho[keep]
    order = np.argsort(rho)
    desc = apply_last_writer(shot_writes(local[order], cosine[order], rho[order], radius), rho.shape[0])
    norm = np.linalg.norm(desc)
    if norm > 0:
        return desc / norm if normalize else desc
    return np.zeros(SHOT_LEN)


def shot_single_scale(
    point_cloud,
    normals,
    keypoints,
    radius: float,
    normalize: bool = True,
    min_neighborhood_size: int = 100,
    return_lrf: bool = False,
):
    """
    shot_parallelization.py:135-183 with `subsampling_voxel_size=None`: KDTree over the cloud, query_radius on
    the keypoint COORDINATES, LRF on all returned neighbours, then the descriptor. Serial (one process).
    """
    neighborhoods = KDTree(point_cloud).query_radius(keypoints, radius)
    out = np.zeros((keypoints.shape[0], SHOT_LEN))
    lrfs = np.zeros((keypoints.shape[0], 3, 3))
    for i, kp in enumerate(keypoints):
        nb = neighborhoods[i]
        lrfs[i] = local_reference_frame(kp, point_cloud[nb], radius)
        out[i] = shot_descriptor(
            kp, point_cloud[nb], normals[nb], radius, lrfs[i], normalize, min_neighborhood_size
        )
    return (out, lrfs) if return_lrf else out


def _shot_task(args):
    kp, pts, nrm, radius, normalize, min_nb = args
    lrf = local_reference_frame(kp, pts, radius)
    return shot_descriptor(kp, pts, nrm, radius, lrf, normalize, min_nb)


def shot_single_scale_pool(
    point_cloud, normals, keypoints, radius: float, normalize: bool, min_neighborhood_size: int, n_procs: int
):
    """
    Same result as `shot_single_scale`, fanned over a `multiprocessing.Pool` the way the reference does
    (shot_parallelization.py:31, :68, :112: the parent gathers one tuple of arrays per query and the workers
    run the per-query functions). Used only as the timed CPU baseline of bench.py.
    """
    from multiprocessing import Pool

    neighborhoods = KDTree(point_cloud).query_radius(keypoints, radius)
    tasks = (
        (kp, point_cloud[neighborhoods[i]], normals[neighborhoods[i]], radius, normalize, min_neighborhood_size)
        for i, kp in enumerate(keypoints)
    )
    chunk = int(np.ceil(keypoints.shape[0] / (2 * n_procs)))
    with Pool(processes=n_procs) as pool:
        rows = list(pool.imap(_shot_task, tasks, chunksize=max(1, min(chunk, 512))))
    return np.array(rows)


def shot_serial_debug(keypoints, cloud_points, normals, radius: float, min_neighborhood_size: int = 10):
    """
    shot.py:310-499, the serial twin: neighbours at distance 0 are dropped BEFORE the LRF (shot.py:361-363),
    the row is always normalised, and the default threshold is 10.
    """
    neighborhoods = KDTree(cloud_points).query_radius(keypoints, radius)
    out = np.zeros((keypoints.shape[0], SHOT_LEN))
    for i, kp in enumerate(keypoints):
        nb = neighborhoods[i]
        pts = cloud_points[nb]
        rho = np.linalg.norm(pts - kp, axis=1)
        keep = rho > 0
        if keep.sum() > min_neighborhood_size:
            lrf = local_reference_frame(kp, pts[keep], radius)
            out[i] = shot_descriptor(kp, pts[keep], normals[nb][keep], radius, lrf, True, min_neighborhood_size)
    return out

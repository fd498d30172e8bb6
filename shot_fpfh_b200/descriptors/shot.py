"""
The function-level API of the reference's `shot_fpfh.descriptors.shot` (shot.py), on the GPU.

`get_local_rf` and `compute_single_shot_descriptor` keep the reference's tuple-argument signatures (they were
`multiprocessing.Pool` task functions, shot.py:16-18, :175-185) and run ONE query through the same kernels as
the batch path; `compute_shot_descriptor` is the serial debug twin (shot.py:310-499) whose semantics differ from
`ShotMultiprocessor` in three ways: neighbours at distance 0 are dropped BEFORE the local reference frame, the
rows are always normalised, and `min_neighborhood_size` defaults to 10.
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt
import torch

from .. import ops
from ..device import Grid, download, upload
from ..neighbors import explicit_csr


def _one_query_grid(neighbors, normals, radius):
    pts = upload(np.asarray(neighbors, dtype=np.float64).reshape(-1, 3))
    nrm = upload(np.asarray(normals, dtype=np.float64).reshape(-1, 3)) if normals is not None else None
    grid = Grid().build(pts, nrm, radius)
    return grid, pts


def get_local_rf(
    values: tuple[npt.NDArray[np.float64], npt.NDArray[np.float64], float],
) -> npt.NDArray[np.float64]:
    """(point, neighbors, radius) -> (3, 3) frame with columns [x y z]; identity without neighbours (shot.py:16-48)."""
    point, neighbors, radius = values
    neighbors = np.asarray(neighbors, dtype=np.float64).reshape(-1, 3)
    if neighbors.shape[0] == 0:
        return np.eye(3)
    grid, pts = _one_query_grid(neighbors, None, radius)
    offsets, nbr = explicit_csr(neighbors.shape[0], pts.device)
    q = upload(np.asarray(point, dtype=np.float64).reshape(1, 3))
    lrf = ops.shot_lrf(grid, q, float(radius), offsets, nbr)[0].cpu().numpy()
    grid.close()
    return lrf


def get_azimuth_idx(x, y):
    """
    Azimuth octant (shot.py:51-70): bins are counted from angle -pi, decided by comparisons only, so that a point
    on a boundary falls in the lower bin. Host-side helper kept for API parity (the kernels use sf::azimuth_octant).
    """
    x, y = np.asarray(x), np.asarray(y)
    upper = (y > 0) | ((y == 0) & (x < 0))
    right = (x > 0) | ((x == 0) & (y > 0))
    inner = np.where((x * y > 0) | (x == 0), np.abs(x) < np.abs(y), np.abs(x) > np.abs(y))
    return 4 * upper + 2 * np.logical_xor(right, upper) + inner


def compute_single_shot_descriptor(values) -> npt.NDArray[np.float64]:
    """
    (point, neighbors, normals, radius, local_rf, normalize, min_neighborhood_size) -> (352,) float64
    (shot.py:175-306).
    """
    point, neighbors, normals, radius, local_rf, normalize, min_neighborhood_size = values
    neighbors = np.asarray(neighbors, dtype=np.float64).reshape(-1, 3)
    if neighbors.shape[0] == 0:
        return np.zeros(ops.SHOT_LEN)
    grid, pts = _one_query_grid(neighbors, normals, radius)
    offsets, nbr = explicit_csr(neighbors.shape[0], pts.device)
    q = upload(np.asarray(point, dtype=np.float64).reshape(1, 3))
    lrf = upload(np.asarray(local_rf, dtype=np.float64).reshape(1, 3, 3))
    desc = ops.shot_descriptor(grid, q, float(radius), offsets, nbr, lrf, int(min_neighborhood_size), bool(normalize))
    out = desc[0].cpu().numpy()
    grid.close()
    return out


def compute_shot_descriptor(
    keypoints: npt.NDArray[np.float64],
    cloud_points: npt.NDArray[np.float64],
    normals: npt.NDArray[np.float64],
    radius: float,
    min_neighborhood_size: int = 10,
    n_cosine_bins: int = 11,
    n_azimuth_bins: int = 8,
    n_elevation_bins: int = 2,
    n_radial_bins: int = 2,
    debug_mode: bool = False,
    disable_progress_bars: bool = True,
) -> npt.NDArray[np.float64]:
    """
    Serial-driver twin (shot.py:310-499). The kernels implement the 11 x 8 x 2 x 2 layout only; the reference
    asserts the last three (shot.py:330-338) and its multiprocess path hard-codes all four (shot.py:197).
    """
    assert n_azimuth_bins == 8, "Generic function for other than 8 azimuth divisions not implemented"
    assert n_elevation_bins == 2, "Generic function for other than 2 elevation divisions not implemented"
    assert n_radial_bins == 2, "Generic function for other than 2 radial divisions not implemented"
    assert n_cosine_bins == 11, "The sm_100a kernel implements 11 cosine bins (the reference's multiprocess layout)"
    pts, nrm, kp = upload(cloud_points), upload(normals), upload(keypoints)
    grid = Grid().build(pts, nrm, radius)
    offsets, nbr, _, dist = ops.radius_csr(grid, kp, radius, want_dist=True)
    # drop the distance-0 neighbours BEFORE the frame (shot.py:361-363): compact the CSR on the device
    keep = dist > 0
    counts = torch.zeros(kp.shape[0] + 1, dtype=torch.int64, device=kp.device)
    seg = torch.repeat_interleave(torch.arange(kp.shape[0], device=kp.device), offsets[1:] - offsets[:-1])
    counts[1:] = torch.bincount(seg[keep], minlength=kp.shape[0])
    offsets_pos = torch.cumsum(counts, 0)
    nbr_pos = nbr[keep].contiguous()
    lrf = ops.shot_lrf(grid, kp, radius, offsets_pos, nbr_pos)
    desc = ops.shot_descriptor(grid, kp, radius, offsets_pos, nbr_pos, lrf, int(min_neighborhood_size), True)
    out = download(desc)
    grid.close()
    return out


def interpolate_on_adjacent_husks(distance, radius: float):
    """
    Radial interpolation weights (shot.py:73-118) -> (outer_bin, inner_bin, current_bin). Host-side helper kept
    for API parity; the kernels evaluate the same piecewise-linear functions in sf::shot_record.
    """
    distance = np.asarray(distance, dtype=np.float64)
    half = radius / 2
    inner_bin = np.where((distance > half) & (distance < radius * 3 / 4), (radius * 3 / 4 - distance) / half, 0.0)
    outer_bin = np.where((distance < half) & (distance > radius / 4), (distance - radius / 4) / half, 0.0)
    current_bin = np.where(
        distance < half,
        1 - np.abs(distance - radius / 4) / half,
        np.where(distance > half, 1 - np.abs(distance - radius * 3 / 4) / half, 0.0),
    )
    return outer_bin, inner_bin, current_bin


def interpolate_vertical_volumes(phi, z):
    """Elevation interpolation weights (shot.py:121-171) -> (upper_volume, lower_volume, current_volume)."""
    phi, z = np.asarray(phi, dtype=np.float64), np.asarray(z, dtype=np.float64)
    half_pi = np.pi / 2
    on_equator = np.abs(phi - half_pi) < 1e-10
    upper_volume = np.where(
        ((phi > half_pi) | (on_equator & (z <= 0))) & (phi <= np.pi * 3 / 4), (np.pi * 3 / 4 - phi) / half_pi, 0.0
    )
    lower_volume = np.where(
        ((phi < half_pi) & (~on_equator | (z > 0))) & (phi >= np.pi / 4), (phi - np.pi / 4) / half_pi, 0.0
    )
    current_volume = np.where(
        phi < half_pi, 1 - np.abs(phi - np.pi / 4) / half_pi, 1 - np.abs(phi - np.pi * 3 / 4) / half_pi
    )
    return upper_volume, lower_volume, current_volume

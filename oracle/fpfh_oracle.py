"""
TEST INFRASTRUCTURE ONLY. NumPy restatement of the reference's FPFH path (shot_fpfh/descriptors/fpfh.py:16-117).

Stage 1 (fpfh.py:38-90): for EVERY cloud point i, a simplified point feature histogram over its radius
neighbourhood (self excluded by `dist > 0`, but counted in the divisor K_i). Stage 2 (fpfh.py:97-116): on
the keypoints (given as INDICES into the cloud), `fpfh = spfh[i] + (sum_{j, d_j > 0} spfh[j] / d_j) / K_i`.

Two layouts:
  * correlated (reference default, runs unmodified): joint n x n x n histogram, flat `(ia*n + ip)*n + it`;
  * decorrelated (3*n bins, the "(N,33)" layout of BASELINE.json): the reference raises on it;
    the layout restated here is the one the one-token patch of oracle/reference_harness.py produces,
    `[alpha bins | phi bins | theta bins]`, each feature binned (and dropped) independently.
Binning is NumPy's: n equal bins on [lo, hi] with edges `linspace(lo, hi, n + 1)`, a value on an interior
edge goes up, `hi` belongs to the last bin, anything outside is dropped.
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt
from sklearn.neighbors import KDTree

RANGES = ((-1.0, 1.0), (-1.0, 1.0), (-np.pi / 2, np.pi / 2))


def bin_edges(n_bins: int) -> npt.NDArray[np.float64]:
    """(3, n_bins + 1) float64 edges exactly as np.histogram / np.histogramdd build them."""
    return np.stack([np.linspace(lo, hi, n_bins + 1) for lo, hi in RANGES])


def bin_index(values: npt.NDArray[np.float64], edges: npt.NDArray[np.float64]) -> npt.NDArray[np.int64]:
    """Bin of each value, -1 when outside [edges[0], edges[-1]]."""
    n_bins = edges.shape[0] - 1
    idx = np.searchsorted(edges, values, side="right") - 1
    idx[values == edges[-1]] = n_bins - 1
    idx[(values < edges[0]) | (values > edges[-1])] = -1
    return idx


def pair_features(point, normal, neighbors, neighbor_normals):
    """fpfh.py:47-57 for the neighbours at distance > 0 -> (alpha, phi, theta)."""
    rel = neighbors - point
    dist = np.linalg.norm(rel, axis=1)
    far = dist > 0
    rel, dist, nn = rel[far], dist[far], neighbor_normals[far]
    v = np.cross(rel, normal)
    w = np.cross(normal, v)
    alpha = np.einsum("ij,ij->i", v, nn)
    phi = rel.dot(normal) / dist
    theta = np.arctan2(np.einsum("ij,ij->i", nn, w), nn.dot(normal))
    return alpha, phi, theta


def spfh_row(alpha, phi, theta, n_neighbors: int, n_bins: int, decorrelated: bool, edges):
    """Integer counts divided by the neighbourhood size INCLUDING the point itself (fpfh.py:79, :88)."""
    ia, ip, it = bin_index(alpha, edges[0]), bin_index(phi, edges[1]), bin_index(theta, edges[2])
    if decorrelated:
        row = np.zeros(3 * n_bins)
        for k, idx in enumerate((ia, ip, it)):
            np.add.at(row, k * n_bins + idx[idx >= 0], 1.0)
    else:
        row = np.zeros(n_bins**3)
        ok = (ia >= 0) & (ip >= 0) & (it >= 0)
        np.add.at(row, (ia[ok] * n_bins + ip[ok]) * n_bins + it[ok], 1.0)
    return row / n_neighbors


def fpfh(
    keypoints_indices,
    cloud_points,
    normals,
    radius: float,
    n_bins: int,
    decorrelated: bool = False,
    return_spfh: bool = False,
):
    """fpfh.py:16-117."""
    neighborhoods, distances = KDTree(cloud_points).query_radius(cloud_points, radius, return_distance=True)
    edges = bin_edges(n_bins)
    width = 3 * n_bins if decorrelated else n_bins**3
    spfh = np.zeros((cloud_points.shape[0], width))
    for i in range(cloud_points.shape[0]):
        nb = neighborhoods[i]
        if nb.shape[0] > 0:
            a, p, t = pair_features(cloud_points[i], normals[i], cloud_points[nb], normals[nb])
            spfh[i] = spfh_row(a, p, t, nb.shape[0], n_bins, decorrelated, edges)
    out = np.zeros((len(keypoints_indices), width))
    for row, i in enumerate(keypoints_indices):
        nb, d = neighborhoods[i], distances[i]
        far = d > 0
        out[row] = spfh[i] + (spfh[nb[far]] / d[far][:, None]).sum(axis=0) / nb.shape[0]
    return (out, spfh) if return_spfh else out

"""
TEST INFRASTRUCTURE ONLY. Fixed-radius neighbour search as the reference performs it.

The reference delegates to scikit-learn's `KDTree(points).query_radius(queries, r[, return_distance=True])`
(call sites: shot_parallelization.py:167-169, :220-222, :229-231, :283-285; fpfh.py:26-30; shot.py:340-341).
scikit-learn is not vendored in the reference (pinned 1.5.1, poetry.lock:916-917). Its published predicate,
read from `sklearn/neighbors/_binary_tree.pxi.tp` (leaf test `rdist <= r**2`, rdist accumulated sequentially
in float64 as `d += tmp * tmp`; returned distance = sqrt(rdist)), is restated in `brute_force_radius`.
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt
from sklearn.neighbors import KDTree


def kdtree_radius(
    support: npt.NDArray[np.float64],
    queries: npt.NDArray[np.float64],
    radius: float,
    return_distance: bool = False,
):
    """Exactly the call the reference makes (tree-traversal order, object arrays)."""
    return KDTree(support).query_radius(queries, radius, return_distance=return_distance)


def brute_force_radius(
    support: npt.NDArray[np.float64],
    queries: npt.NDArray[np.float64],
    radius: float,
) -> tuple[npt.NDArray[np.int64], npt.NDArray[np.int64], npt.NDArray[np.float64]]:
    """
    The predicate without the tree, as CSR (offsets, indices ascending per query, distances).
    `((dx*dx + dy*dy) + dz*dz) <= r*r` in float64, inclusive, no fused multiply-add.
    """
    r2 = np.float64(radius) * np.float64(radius)
    offsets = np.zeros(queries.shape[0] + 1, dtype=np.int64)
    idx_chunks, dist_chunks = [], []
    for i, q in enumerate(queries):
        d = support - q
        sq = d * d
        rdist = (sq[:, 0] + sq[:, 1]) + sq[:, 2]
        hit = np.nonzero(rdist <= r2)[0]
        idx_chunks.append(hit)
        dist_chunks.append(np.sqrt(rdist[hit]))
        offsets[i + 1] = offsets[i] + hit.shape[0]
    indices = np.concatenate(idx_chunks) if idx_chunks else np.zeros(0, dtype=np.int64)
    dists = np.concatenate(dist_chunks) if dist_chunks else np.zeros(0)
    return offsets, indices.astype(np.int64), dists


def to_sorted_csr(neighborhoods, distances=None):
    """Object array of index arrays (tree order) -> CSR with ascending indices per query."""
    q = len(neighborhoods)
    offsets = np.zeros(q + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(n) for n in neighborhoods])
    indices = np.zeros(offsets[-1], dtype=np.int64)
    dists = np.zeros(offsets[-1]) if distances is not None else None
    for i in range(q):
        order = np.argsort(neighborhoods[i], kind="stable")
        indices[offsets[i] : offsets[i + 1]] = neighborhoods[i][order]
        if distances is not None:
            dists[offsets[i] : offsets[i + 1]] = distances[i][order]
    return offsets, indices, dists

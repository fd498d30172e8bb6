"""
`RadiusSearch`: the drop-in for `sklearn.neighbors.KDTree(points).query_radius(queries, r)` as the reference uses
it (shot_parallelization.py:167-169, fpfh.py:26-30, shot.py:340-341), backed by the uniform grid of csrc/grid.cu.

Same results as the KD-tree as SETS (bit-exact float64 predicate); the order inside each neighbourhood is the
grid walk order instead of the tree traversal order (the reference never relies on that order, except through
the unstable `np.argsort(rho)` on exact distance ties — DESIGN.md "Ties").
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt
import torch

from . import ops
from .device import Grid, upload


class RadiusSearch:
    """`RadiusSearch(points, radius)` builds the grid once; `query_radius` mirrors KDTree.query_radius."""

    def __init__(self, points, radius: float, normals=None) -> None:
        self.points_dev = upload(points)
        self.normals_dev = upload(normals) if normals is not None else None
        self.grid = Grid().build(self.points_dev, self.normals_dev, radius)
        self.radius = float(radius)

    def csr(self, queries, radius: float | None = None, want_index=True, want_dist=False, want_sorted=False):
        """Device CSR: (offsets, nbr_sorted, nbr_index, dist) — see ops.radius_csr."""
        q = None if queries is None else upload(queries)
        return ops.radius_csr(
            self.grid, q, self.radius if radius is None else radius, want_sorted=want_sorted, want_index=want_index,
            want_dist=want_dist,
        )

    def query_radius(
        self, queries: npt.NDArray[np.float64], r: float | None = None, return_distance: bool = False
    ):
        """Object array of int64 index arrays (and of float64 distance arrays), like sklearn."""
        offsets, _, index, dist = self.csr(queries, r, want_index=True, want_dist=return_distance)
        offs = offsets.cpu().numpy()
        idx = index.cpu().numpy().astype(np.int64)
        nq = offs.shape[0] - 1
        out = np.empty(nq, dtype=object)
        for i in range(nq):
            out[i] = idx[offs[i] : offs[i + 1]]
        if not return_distance:
            return out
        d = dist.cpu().numpy()
        dout = np.empty(nq, dtype=object)
        for i in range(nq):
            dout[i] = d[offs[i] : offs[i + 1]]
        return out, dout

    def close(self) -> None:
        self.grid.close()


def explicit_csr(n_neighbors: int, device: torch.device):
    """CSR of ONE query whose neighbourhood is every point of the grid (cell-sorted positions 0..n-1)."""
    offsets = torch.tensor([0, n_neighbors], dtype=torch.int64, device=device)
    nbr = torch.arange(n_neighbors, dtype=torch.int32, device=device)
    return offsets, nbr

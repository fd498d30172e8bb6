"""
Voxel-grid subsampling with the semantics of the reference's `grid_subsampling` (core/subsampling.py:5-39):
voxel key = floor((p - min) / voxel) per axis, voxels in lexicographic key order (what `np.unique(axis=0)`
returns), and per voxel the index of the point closest to the voxel's barycentre (first one on ties).

This sits immediately UPSTREAM of the hot path (it only reduces the support cloud when the pipeline passes
`subsampling_voxel_size`, shot_parallelization.py:157-161) and stays on the host in this round (SURVEY.md §8f
ranks its GPU version as the next row). It is vectorised: the reference loops over voxels in Python.
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt


def grid_subsampling_gpu(points: npt.NDArray[np.float64], voxel_size: float) -> npt.NDArray[np.int64]:
    """The same selection on the device (csrc/subsample.cu); what `ShotMultiprocessor` uses for its support."""
    from . import ops
    from .device import upload

    return ops.voxel_subsample(upload(points), voxel_size).cpu().numpy()


def grid_subsampling(points: npt.NDArray[np.float64], voxel_size: float) -> npt.NDArray[np.int64]:
    points = np.asarray(points, dtype=np.float64)
    if points.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    keys = ((points - np.min(points, axis=0)) // voxel_size).astype(np.int64)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))  # stable: ascending index inside a voxel
    sk = keys[order]
    first = np.ones(points.shape[0], dtype=bool)
    first[1:] = np.any(sk[1:] != sk[:-1], axis=1)
    starts = np.nonzero(first)[0]
    counts = np.diff(np.append(starts, points.shape[0]))
    sp = points[order]
    means = np.add.reduceat(sp, starts, axis=0) / counts[:, None]
    group = np.repeat(np.arange(starts.shape[0]), counts)
    dist = np.linalg.norm(sp - means[group], axis=1)
    # argmin per group, first occurrence on ties: sort by (group, dist, position) and take each group's head
    pick = np.lexsort((np.arange(points.shape[0]), dist, group))[starts]
    return order[pick].astype(np.int64)

"""
`compute_fpfh_descriptor` with the reference's signature (shot_fpfh/descriptors/fpfh.py:16-25), computed by the
sm_100a kernels of csrc/grid.cu (radius search over ALL cloud points) and csrc/fpfh.cu (SPFH, then FPFH).

`decorrelated=True` (the 3 * n_bins layout, e.g. (N, 33) for n_bins = 11) RAISES in the reference because a
(n_bins, 3) array is assigned into a (3 * n_bins,) row (fpfh.py:59-79). Here it works and returns
the concatenated layout `[alpha bins | phi bins | theta bins]` — what the reference yields once the single token
`).T` at fpfh.py:78 is replaced by `).ravel()`.
"""

from __future__ import annotations

import logging

import numpy as np
import numpy.typing as npt
import torch

from .. import ops
from ..device import DenseRowsDownload, Grid, remember_device_rows, upload


def fpfh_device(grid: Grid, keypoints_dev: torch.Tensor, radius: float, n_bins: int, decorrelated: bool,
                out_dtype: torch.dtype = torch.float64):
    """
    search over every cloud point -> SPFH (cell-sorted rows) -> FPFH on the keypoints, by the fused driver in ONE
    call (sf_fpfh_cloud), rows left on the device. `compute_fpfh_descriptor` and the multi-GPU driver use its block
    form (ops.FpfhBlock: same kernels, same results, the FPFH stage callable per block of keypoints / per block of the
    cloud). Returns (fpfh, mean K).
    """
    out, pairs = ops.fpfh_cloud(grid, radius, n_bins, decorrelated, keypoints_dev, out_dtype=out_dtype)
    return out, float(pairs) / max(grid.n, 1)


def compute_fpfh_descriptor(
    keypoints_indices: npt.NDArray[np.int64],
    cloud_points: npt.NDArray[np.float64],
    normals: npt.NDArray[np.float64],
    radius: float,
    n_bins: int,
    decorrelated: bool = False,
    verbose: bool = True,
    disable_progress_bars: bool = True,
) -> npt.NDArray[np.float64]:
    """
    FPFH on the points `cloud_points[keypoints_indices]` -> (Q, n_bins**3) float64, or (Q, 3 * n_bins) when
    `decorrelated`. `keypoints_indices` are INDICES into the cloud (pipeline.py:330), unlike SHOT's coordinates.
    """
    pts, nrm = upload(cloud_points), upload(normals)  # asynchronous from pinned memory: issued before the host checks
    kp = np.asarray(keypoints_indices)
    if kp.size:
        lowest, highest = int(kp.min()), int(kp.max())
        if lowest < -cloud_points.shape[0] or highest >= cloud_points.shape[0]:
            raise IndexError("keypoints_indices out of bounds for the point cloud")
        if lowest < 0:  # NumPy's negative indexing, as `cloud_points[keypoints_indices]` would resolve it
            kp = np.where(kp < 0, kp + cloud_points.shape[0], kp)
    grid = _cached_grid().build(pts, nrm, radius)
    kp_dev = upload(kp, torch.int64)
    n_kp = int(kp_dev.shape[0])
    width = 3 * int(n_bins) if decorrelated else int(n_bins) ** 3
    if n_kp == 0:
        return np.zeros((0, width))
    # The fused driver in its block form (one block = the whole cloud): one scan of the candidates + SPFH of every
    # point, then the FPFH rows by blocks of keypoints, so that the float32 rows of block b cross PCIe (and are
    # widened exactly to the float64 array the reference returns by host threads) while block b + 1 is computed.
    block = ops.FpfhBlock(grid, float(radius), int(n_bins), bool(decorrelated), 0, grid.n, pts.device)
    spfh_all = block.spfh(want_pairs=verbose)
    if verbose:
        logging.info(f"Mean neighborhood size over the whole point cloud: {block.pairs / max(grid.n, 1):.2f}")
    parts = max(1, min(_OUTPUT_BLOCKS, n_kp // 32768))
    job = DenseRowsDownload(n_kp, width)
    pieces = []
    for b in range(parts):
        lo, hi = n_kp * b // parts, n_kp * (b + 1) // parts
        pieces.append(block.rows(spfh_all, kp_dev[lo:hi].contiguous(), out_dtype=torch.float32))
        job.push(pieces[-1])
    result = job.finish()
    remember_device_rows(result, pieces[0] if len(pieces) == 1 else torch.cat(pieces))  # hand-off to the matcher
    return result


_OUTPUT_BLOCKS = 8
_GRIDS: dict[int, Grid] = {}


def _cached_grid() -> Grid:
    """One grid handle per device, kept between calls: its buffers (72 bytes per point) are reused instead of being
    allocated and freed by every call (a dozen cudaMalloc/cudaFree, several milliseconds). `release_cached_grids()`
    gives the memory back."""
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    if dev not in _GRIDS:
        _GRIDS[dev] = Grid()
    return _GRIDS[dev]


def release_cached_grids() -> None:
    for grid in _GRIDS.values():
        grid.close()
    _GRIDS.clear()

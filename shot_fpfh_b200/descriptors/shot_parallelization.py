"""
`ShotMultiprocessor` — same dataclass fields, context-manager protocol, method names, argument meaning and outputs
as the reference class (shot_fpfh/descriptors/shot_parallelization.py:16-312), computed by the sm_100a kernels of
csrc/grid.cu and csrc/shot.cu instead of a `multiprocessing.Pool` of NumPy workers.

    with ShotMultiprocessor(normalize=True, min_neighborhood_size=10) as shot:
        desc = shot.compute_descriptor_single_scale(point_cloud, normals, keypoints, radius)   # (Q, 352) float64

`n_procs` is accepted and ignored (the parallelism is the GPU's), `disable_progress_bar` likewise (a launch takes
milliseconds). All entry points synchronise before returning host NumPy arrays.
"""

from __future__ import annotations

import logging
from dataclasses import dataclass
from types import TracebackType

import numpy as np
import numpy.typing as npt
import torch

from .. import ops
from ..device import (Grid, SparseRowsDownload, download, download_sparse_rows, remember_device_rows, require_cuda,
                      upload)
from ..distributed import block_bounds


def _neighborhoods_to_csr(neighborhoods, inv_perm: torch.Tensor):
    """Object array of index arrays (as KDTree.query_radius returns) -> device CSR of cell-sorted positions."""
    counts = np.fromiter((len(n) for n in neighborhoods), dtype=np.int64, count=len(neighborhoods))
    offsets = np.zeros(len(neighborhoods) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    flat = np.concatenate([np.asarray(n, dtype=np.int64) for n in neighborhoods]) if offsets[-1] else np.zeros(0, np.int64)
    idx = upload(flat, torch.int64)
    return upload(offsets, torch.int64), inv_perm[idx].contiguous()


@dataclass
class ShotMultiprocessor:
    """
    Base class to compute SHOT descriptors (reference: shot_parallelization.py:16-28; defaults identical,
    including `min_neighborhood_size=100`, which zeroes every descriptor on sparse clouds — SURVEY.md F4).
    """

    normalize: bool = True
    share_local_rfs: bool = True
    min_neighborhood_size: int = 100

    n_procs: int = 8
    disable_progress_bar: bool = False
    verbose: bool = True

    def __enter__(self):
        require_cuda()  # fails loudly here rather than at the first kernel
        self._grid = None
        self._ensure_grid()
        return self

    def __exit__(
        self,
        exc_type: type | None,
        exc_val: Exception | None,
        exc_tb: TracebackType | None,
    ) -> None:
        torch.cuda.synchronize()
        grid = getattr(self, "_grid", None)
        if grid is not None:
            grid.close()
            self._grid = None

    # ------------------------------------------------------------------------------------------------------
    def _ensure_grid(self) -> Grid:
        if getattr(self, "_grid", None) is None:
            # blocks of queries after the first size their neighbour lists without a host round trip; `poll` below
            self._grid = Grid().set_speculative(builds=False, shot_lists=True)
        return self._grid

    def _support(self, point_cloud, normals, subsampling_voxel_size, radius):
        """
        Uploads the cloud, reduces it to the voxel-subsampled support when asked (on the device: csrc/subsample.cu,
        `grid_subsampling` semantics) and builds the grid for `radius`.
        """
        pts, nrm = upload(point_cloud), upload(normals)
        if subsampling_voxel_size is not None:
            support = ops.voxel_subsample(pts, subsampling_voxel_size)
            if self.verbose:
                logging.info(
                    f"Keeping a support of {support.shape[0]} points out of {pts.shape[0]} "
                    f"(voxel size: {subsampling_voxel_size:.2f})"
                )
            pts, nrm = pts[support].contiguous(), nrm[support].contiguous()
        grid = self._ensure_grid().build(pts, nrm, radius)
        return grid, pts, nrm

    def _to_host(self, t: torch.Tensor) -> npt.NDArray[np.float64]:
        """float64 host array; float32 SHOT rows cross PCIe compacted (device.SparseRowsDownload)."""
        if t.dtype == torch.float32:
            return download_sparse_rows(t, self.n_procs)
        return download(t.double())

    # queries per block of the pipelined single-scale call: block i's rows are rebuilt on the host while block
    # i + 1 is computed and copied
    _PIPELINE_BLOCK = 16384
    _PIPELINE_MAX_BLOCKS = 4

    # ------------------------------------------------------------------------------------------------------
    def compute_local_rf(
        self,
        keypoints: npt.NDArray[np.float64],
        neighborhoods: np.ndarray,
        support: npt.NDArray[np.float64],
        radius: float,
    ) -> npt.NDArray[np.float64]:
        """
        Local reference frames of the keypoints from caller-provided neighbourhoods (index arrays into
        `support`), as reference `compute_local_rf` (shot_parallelization.py:46-84) -> (Q, 3, 3) float64.
        """
        pts = upload(support)
        grid = self._ensure_grid().build(pts, None, radius)
        _, inv_perm = ops.grid_permutation(grid)
        offsets, nbr = _neighborhoods_to_csr(neighborhoods, inv_perm)
        return self._to_host(ops.shot_lrf(grid, upload(keypoints), radius, offsets, nbr))

    def compute_descriptor(
        self,
        keypoints: npt.NDArray[np.float64],
        normals: npt.NDArray[np.float64],
        neighborhoods: np.ndarray,
        local_rfs: npt.NDArray[np.float64],
        support: npt.NDArray[np.float64],
        radius: float,
    ) -> npt.NDArray[np.float64]:
        """
        SHOT descriptors from caller-provided neighbourhoods and frames, as reference `compute_descriptor`
        (shot_parallelization.py:86-133) -> (Q, 352) float64.
        """
        pts, nrm = upload(support), upload(normals)
        grid = self._ensure_grid().build(pts, nrm, radius)
        _, inv_perm = ops.grid_permutation(grid)
        offsets, nbr = _neighborhoods_to_csr(neighborhoods, inv_perm)
        desc = ops.shot_descriptor(
            grid, upload(keypoints), radius, offsets, nbr, upload(local_rfs), self.min_neighborhood_size, self.normalize
        )
        return self._to_host(desc)

    # ------------------------------------------------------------------------------------------------------
    def _single_scale_device(self, grid, keypoints_dev, lrf_radius, shot_radius, lrf=None, out_dtype=torch.float64,
                             need_lrf=False):
        """search -> LRF -> descriptor, all on the device. Returns (descriptors, lrf)."""
        if lrf is None and shot_radius == lrf_radius:
            # frames and descriptors on the same neighbourhoods: the fused driver (one pass over the candidate cells)
            desc, lrf, _ = ops.shot_single_scale(
                grid, keypoints_dev, shot_radius, self.min_neighborhood_size, self.normalize, out_dtype=out_dtype,
                want_lrf=need_lrf,
            )
            return desc, lrf
        offsets, nbr, _, _ = ops.radius_csr(grid, keypoints_dev, lrf_radius)
        if lrf is None:
            lrf = ops.shot_lrf(grid, keypoints_dev, lrf_radius, offsets, nbr)
        if shot_radius != lrf_radius:
            offsets, nbr, _, _ = ops.radius_csr(grid, keypoints_dev, shot_radius)
        desc = ops.shot_descriptor(
            grid, keypoints_dev, shot_radius, offsets, nbr, lrf, self.min_neighborhood_size, self.normalize,
            out_dtype=out_dtype,
        )
        return desc, lrf

    def compute_descriptor_single_scale(
        self,
        point_cloud: npt.NDArray[np.float64],
        normals: npt.NDArray[np.float64],
        keypoints: npt.NDArray[np.float64],
        radius: float,
        subsampling_voxel_size: float | None = None,
    ) -> npt.NDArray[np.float64]:
        """
        SHOT on a single scale (reference: shot_parallelization.py:135-183). `keypoints` are COORDINATES (Q, 3).
        Returns the descriptors as a (Q, 352) float64 array.
        """
        n_kp = int(np.shape(keypoints)[0])
        if n_kp == 0:
            return np.zeros((0, 352))
        for attempt in range(2):
            # Here the reference's `n_procs` workers are the host threads that rebuild the dense float64 rows.
            job = SparseRowsDownload(n_kp, 352, self.n_procs)
            try:
                self._ensure_grid().set_speculative(builds=False, shot_lists=attempt == 0)
                grid, _, _ = self._support(point_cloud, normals, subsampling_voxel_size, radius)
                kp = upload(keypoints)
                blocks = max(1, min(self._PIPELINE_MAX_BLOCKS, n_kp // self._PIPELINE_BLOCK))
                rows = torch.empty((n_kp, 352), dtype=torch.float32, device=kp.device)  # kept for the matcher (hand-off)
                for b in range(blocks):
                    lo, hi = block_bounds(n_kp, blocks, b)
                    ops.shot_single_scale(grid, kp[lo:hi], radius, self.min_neighborhood_size, self.normalize,
                                          out=rows[lo:hi])
                    job.push(rows[lo:hi])
                result = job.finish()
                self.last_d2h_bytes = job.bytes_copied  # read by bench.py
            finally:
                job.abandon()
            # blocks after the first sized their neighbour lists from the first one's (no host round trip); when a
            # block needed more, its kernels did nothing and said so: the call is repeated with exact sizes
            if grid.poll() == 0:
                remember_device_rows(result, rows)
                return result
        raise RuntimeError("SHOT: the device-side size check failed on a synchronising call")  # cannot happen

    def compute_descriptor_bi_scale(
        self,
        point_cloud: npt.NDArray[np.float64],
        normals: npt.NDArray[np.float64],
        keypoints: npt.NDArray[np.float64],
        local_rf_radius: float,
        shot_radius: float,
        subsampling_voxel_size: float | None = None,
    ) -> npt.NDArray[np.float64]:
        """
        Two radii: one for the local reference frames, one for the descriptor (reference:
        shot_parallelization.py:185-239). The reference crashes when `subsampling_voxel_size` is None
        (it indexes `point_cloud[None]` at :229); here None keeps the whole support.
        """
        grid, _, _ = self._support(point_cloud, normals, subsampling_voxel_size, max(local_rf_radius, shot_radius))
        desc, _ = self._single_scale_device(
            grid, upload(keypoints), local_rf_radius, shot_radius, out_dtype=torch.float32
        )
        return self._to_host(desc)

    def compute_descriptor_multiscale(
        self,
        point_cloud: npt.NDArray[np.float64],
        normals: npt.NDArray[np.float64],
        keypoints: npt.NDArray[np.float64],
        radii: list[float] | npt.NDArray[np.float64],
        voxel_sizes: list[float] | npt.NDArray[np.float64] | None = None,
        weights: list[float] | npt.NDArray[np.float64] | None = None,
    ) -> npt.NDArray[np.float64]:
        """
        SHOT on several scales (reference: shot_parallelization.py:241-312). With `share_local_rfs` the frames of
        the first radius are reused. The (n_scales, Q, 352) stack is reshaped to (Q, 352 * n_scales) WITHOUT a
        transpose, exactly as the reference does at :312 — so a row is not one keypoint's
        scales side by side; kept for output parity.
        """
        if weights is None:
            weights = np.ones(len(radii))
        kp = upload(keypoints)
        all_descriptors = np.zeros((len(radii), keypoints.shape[0], 352))
        lrf = None
        for scale, radius in enumerate(radii):
            radius = float(radius)
            voxel = None if voxel_sizes is None else float(voxel_sizes[scale])
            grid, _, _ = self._support(point_cloud, normals, voxel, radius)
            desc, new_lrf = self._single_scale_device(
                grid, kp, radius, radius, lrf=lrf if self.share_local_rfs else None, need_lrf=self.share_local_rfs,
                out_dtype=torch.float32,
            )
            if lrf is None or not self.share_local_rfs:
                lrf = new_lrf
            all_descriptors[scale] = self._to_host(desc) * weights[scale]
        return all_descriptors.reshape(keypoints.shape[0], 352 * len(radii))

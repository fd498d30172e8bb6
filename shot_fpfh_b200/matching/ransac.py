"""
`ransac_on_matches` with the reference's signature and random stream (shot_fpfh/matching/ransac.py:14-82).

What the reference does per draw: pick `draw_size` matches with its module-level `default_rng(72)`, fit a rigid
transform to them (Kabsch), count the matches the transform brings within `distance_threshold`; keep the first draw
with the largest count. The cost is the n_draws x n_matches count: that runs on the device
(`sf_ransac_count_inliers`, csrc/registration.cu). The draws replay NumPy's generator on the host — the same
generator, seeded and advanced the same way, so the same draws — and the n_draws 3x3 SVDs are one batched host call.
The winning transform is then refitted with the scalar solver, i.e. exactly the reference's arithmetic.
"""

from __future__ import annotations

import logging

import numpy as np
import numpy.typing as npt

from ..core import RigidTransform, solver_point_to_point

# the reference's seed and its module-level lifetime (ransac.py:14): successive calls continue the stream
rng = np.random.default_rng(seed=72)


def _batched_kabsch(scan: npt.NDArray[np.float64], ref: npt.NDArray[np.float64]) -> npt.NDArray[np.float64]:
    """(D, k, 3) x2 -> (D, 12) rows [rotation row-major | translation] of the fit of each draw (solvers.py:9-31)."""
    scan_centre, ref_centre = scan.mean(axis=1), ref.mean(axis=1)
    cov = np.einsum("dki,dkj->dij", scan - scan_centre[:, None, :], ref - ref_centre[:, None, :])
    u, _, vt = np.linalg.svd(cov)
    rotation = np.einsum("dki,djk->dij", vt, u)  # V U^T
    reflected = np.linalg.det(rotation) < 0
    if reflected.any():
        u = u.copy()
        u[reflected, :, -1] *= -1  # the reference negates the last row of U^T
        rotation = np.einsum("dki,djk->dij", vt, u)
    translation = ref_centre - np.einsum("dij,dj->di", rotation, scan_centre)
    return np.concatenate([rotation.reshape(-1, 9), translation], axis=1)


def ransac_on_matches(
    scan_descriptors_indices: np.ndarray,
    ref_descriptors_indices: np.ndarray,
    scan_keypoints: npt.NDArray[np.float64],
    ref_keypoints: npt.NDArray[np.float64],
    n_draws: int = 10000,
    draw_size: int = 4,
    distance_threshold: float = 1,
    verbose: bool = False,
    disable_progress_bar: bool = False,
) -> tuple[float, RigidTransform]:
    """Returns (share of the matches that are inliers of the best draw, its rigid transform), as the reference."""
    import torch

    from .. import ops
    from ..device import upload

    n_matches = scan_descriptors_indices.shape[0]
    matched_scan = np.ascontiguousarray(scan_keypoints[scan_descriptors_indices], dtype=np.float64)
    matched_ref = np.ascontiguousarray(ref_keypoints[ref_descriptors_indices], dtype=np.float64)
    # ransac.py:48-53, one call per draw: the generator's stream is what makes the result reproducible
    draws = np.stack([rng.choice(n_matches, draw_size, replace=False, shuffle=False) for _ in range(n_draws)])
    transforms = _batched_kabsch(matched_scan[draws], matched_ref[draws])
    counts = ops.ransac_count_inliers(upload(matched_scan), upload(matched_ref), upload(transforms),
                                      float(distance_threshold))
    best = int(torch.argmax(counts).item())  # first maximum = the reference's strict `>` update (ransac.py:63)
    best_n_inliers = int(counts[best].item())
    if verbose:
        logging.info(f"Best draw: #{best} with {best_n_inliers} inliers out of {n_matches} matches")
    best_transform = solver_point_to_point(matched_scan[draws[best]], matched_ref[draws[best]])
    best_transform.normalize_rotation()
    return best_n_inliers / n_matches, best_transform

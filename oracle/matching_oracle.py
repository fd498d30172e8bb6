"""
TEST INFRASTRUCTURE ONLY. NumPy/SciPy restatement of the reference's descriptor matching
(shot_fpfh/matching/matching.py:9-146, :149-169, :172-221 and matching/filters.py:19-40).

`scipy.spatial.distance.cdist` (Euclidean, float64: `sqrt(sum((a-b)^2))` accumulated sequentially) is a
third-party native the reference calls and does not vendor (pinned 1.14.0, poetry.lock:962-963).
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt
from scipy.spatial.distance import cdist


def nonempty_rows(descriptors: npt.NDArray[np.float64]) -> npt.NDArray[np.int64]:
    """matching.py:43-44 / :162-163: rows with at least one non-zero entry."""
    return np.any(descriptors, axis=1).nonzero()[0]


def nearest(scan, ref):
    """
    -> (scan_ids, ref_ids, argmin over the non-empty ref rows, nn distance, distance matrix).
    argmin takes the lowest index on exact ties (NumPy).
    """
    scan_ids, ref_ids = nonempty_rows(scan), nonempty_rows(ref)
    dmat = cdist(scan[scan_ids], ref[ref_ids])
    nn = dmat.argmin(axis=1)
    return scan_ids, ref_ids, nn, dmat[np.arange(scan_ids.shape[0]), nn], dmat


def basic_matching(scan, ref):
    """matching.py:149-169."""
    scan_ids, ref_ids, nn, _, _ = nearest(scan, ref)
    return scan_ids, ref_ids[nn]


def threshold_filter(distances, threshold_multiplier: float):
    """filters.py:19-23."""
    return distances <= distances[distances.nonzero()[0]].min() * threshold_multiplier


def quantile_filter(distances, quantiles):
    """filters.py:26-31."""
    lo, hi = np.quantile(distances, quantiles)
    return (distances >= lo) & (distances <= hi)


def match_descriptors(scan, ref, filter_callback=None, filter_nonreciprocal=False, n_min_matches=100, **kwargs):
    """matching.py:39-74 and :143-146 (the 2-D branch, the only one the pipeline reaches)."""
    scan_ids, ref_ids, nn, dist, dmat = nearest(scan, ref)
    mask = filter_callback(dist, **kwargs) if filter_callback is not None else np.ones(dist.shape[0], dtype=bool)
    if filter_nonreciprocal:
        reciprocal = dmat.argmin(axis=0)[nn] == np.arange(nn.shape[0])
        both = mask & reciprocal
        if both.sum() >= n_min_matches:
            mask = both
    return scan_ids[mask], ref_ids[nn[mask]]


def match_multiscale(scan, ref, filter_callback=None, n_min_matches=100, filter_nonreciprocal=False, **kwargs):
    """
    matching.py:76-136, the 3-D branch: (n_scales, n_points, width) descriptors, distance = minimum over the scales
    of the per-scale Euclidean distances, 1000 for empty rows; dense matrices as in the reference. The reference's
    reciprocity filter assigns into a temporary (a no-op) and is therefore absent here; its
    "too few matches" fallback re-runs without it, which gives the same selection.
    """
    max_val = 1000
    n_scales, n_points, _ = scan.shape
    n_points_ref = ref.shape[1]
    inf = np.ones((n_points, n_points_ref)) * max_val
    for s in range(n_scales):
        ne_a, ne_b = np.any(scan[s], axis=1), np.any(ref[s], axis=1)
        d = np.ones((n_points, n_points_ref)) * max_val
        d[np.ix_(ne_a, ne_b)] = cdist(scan[s][ne_a], ref[s][ne_b])
        inf = np.minimum(d, inf)
    idx = inf.argmin(axis=1)
    dist = inf[np.arange(n_points), idx]
    mask = (filter_callback(dist, **kwargs) if filter_callback is not None else np.ones(n_points, dtype=bool)) & (
        dist < max_val
    )
    return np.arange(n_points)[mask], np.arange(n_points_ref)[idx[mask]]


def ratio_matching(scan, ref, threshold: float, lowe: bool = False):
    """
    The ratio test matching.py:172-221 INTENDS (the reference raises on every input): nearest
    and second-nearest distances d1 <= d2 per non-empty scan row, ratio d1/d2 (1 where d2 == 0), keep the
    rows with `ratio >= threshold` as coded at matching.py:203-211 (`lowe=True` keeps `ratio < threshold`,
    Lowe's sense). Returns (scan ids kept, ref id of the nearest neighbour of each).
    PARITY UNPINNED BY THE REFERENCE — this is the documented restatement.
    """
    scan_ids, ref_ids, nn, d1, dmat = nearest(scan, ref)
    if dmat.shape[1] < 2:
        d2 = np.zeros_like(d1)
    else:
        d2 = np.partition(dmat, 1, axis=1)[:, 1]
    ratio = np.divide(d1, d2, out=np.ones_like(d1), where=d2 != 0)
    keep = ratio < threshold if lowe else ratio >= threshold
    return scan_ids[keep], ref_ids[nn[keep]]

"""
§8f row 4 on the device: RANSAC's inlier counts and the point-to-plane ICP iteration (csrc/registration.cu) behind
`ransac_on_matches` / `icp_point_to_plane`, against oracle/registration_oracle.py (pinned bit for bit to the
reference in tests/test_oracle_golden.py).
"""

import numpy as np
import pytest
from conftest import registration_case

from oracle import registration_oracle as ro

pytestmark = pytest.mark.gpu


def test_ransac_replays_the_reference_stream_and_counts_exactly():
    import shot_fpfh_b200.matching.ransac as ransac

    scan, ref, _, scan_idx, ref_idx = registration_case()
    for n_draws, threshold in ((300, 0.05), (2000, 0.01)):
        ransac.rng = np.random.default_rng(seed=72)  # the module's state at import time, as the reference's
        ratio, t = ransac.ransac_on_matches(scan_idx, ref_idx, scan, ref, n_draws=n_draws, distance_threshold=threshold,
                                            disable_progress_bar=True)
        want_ratio, want_t = ro.ransac_on_matches(scan_idx, ref_idx, scan, ref, np.random.default_rng(seed=72),
                                                  n_draws=n_draws, distance_threshold=threshold)
        assert ratio == want_ratio  # integer count / integer: the same draw won with the same number of inliers
        # the winner is refitted with the reference's own scalar arithmetic: identical, not just close
        assert np.array_equal(t.rotation, want_t[0]) and np.array_equal(t.translation, want_t[1])
        assert ratio > 0.5
    # successive calls continue the stream, like the reference's module-level generator
    again = ransac.ransac_on_matches(scan_idx, ref_idx, scan, ref, n_draws=50, distance_threshold=0.05)[1]
    assert not np.array_equal(again.rotation, t.rotation)


def test_ransac_counts_every_draw():
    """The device count of each draw against NumPy on the same transforms."""
    import torch

    from shot_fpfh_b200 import ops
    from shot_fpfh_b200.device import upload
    from shot_fpfh_b200.matching.ransac import _batched_kabsch

    scan, ref, _, scan_idx, ref_idx = registration_case()
    a, b = scan[scan_idx], ref[ref_idx]
    rng = np.random.default_rng(3)
    draws = np.stack([rng.choice(a.shape[0], 4, replace=False) for _ in range(500)])
    transforms = _batched_kabsch(a[draws], b[draws])
    got = ops.ransac_count_inliers(upload(a), upload(b), upload(transforms), 0.03).cpu().numpy()
    want = np.array([
        (np.linalg.norm(a @ t[:9].reshape(3, 3).T + t[9:] - b, axis=1) <= 0.03).sum() for t in transforms
    ])
    assert np.array_equal(got, want) and want.max() > 300
    assert ops.ransac_count_inliers(upload(a), upload(b), upload(np.zeros((0, 12))), 0.03).shape == (0,)
    assert torch.cuda.is_available()


def test_icp_point_to_plane_against_oracle():
    import torch

    from shot_fpfh_b200 import ops, synthetic
    from shot_fpfh_b200.core import RigidTransform
    from shot_fpfh_b200.device import Grid, upload
    from shot_fpfh_b200.icp import icp_point_to_plane

    scan, ref, normals, scan_idx, ref_idx = registration_case()
    s = np.sqrt(3.0 / scan.shape[0])  # three unit faces
    _, init = ro.ransac_on_matches(scan_idx, ref_idx, scan, ref, np.random.default_rng(seed=72), n_draws=300,
                                   distance_threshold=0.05)
    # one iteration's pieces: nearest reference point within d_max == KDTree.query, sums == NumPy's
    from sklearn.neighbors import KDTree

    d_max = 4 * s
    sub = ops.voxel_subsample(upload(scan), 3 * s).cpu().numpy()
    moved = scan[sub] @ init[0].T + init[1]
    dist, nn = KDTree(ref).query(moved)
    keep = dist[:, 0] <= d_max
    grid = Grid().build(upload(ref), upload(normals), d_max)
    row = np.concatenate([init[0].ravel(), init[1]])
    sums, nearest = ops.icp_plane_step(grid, upload(scan[sub]), row, d_max, want_nearest=True)
    grid.close()
    nearest = nearest.cpu().numpy()
    assert np.array_equal(nearest[keep], nn[keep, 0]) and (nearest[~keep] == -1).all() and keep.sum() > 100
    p, q, n = moved[keep], ref[nn[keep, 0]], normals[nn[keep, 0]]
    g = np.hstack((np.cross(p, n), n))
    h = np.einsum("ij,ij->i", q - p, n)
    assert np.allclose(sums[:21], (g.T @ g)[np.triu_indices(6)], rtol=1e-11, atol=1e-13)
    assert np.allclose(sums[21:27], g.T @ h, rtol=1e-9, atol=1e-13)
    assert np.isclose(sums[27], np.abs(h).sum(), rtol=1e-11) and sums[28] == keep.sum()

    # the whole loop, both sides on the same subsampled scan points
    for max_iter, rms_threshold in ((20, 1e-6), (50, 1.6e-3)):  # runs out of iterations / stops early
        t, rms, ok = icp_point_to_plane(scan, ref, normals, RigidTransform(init[0], init[1]), d_max=d_max,
                                        voxel_size=3 * s, max_iter=max_iter, rms_threshold=rms_threshold)
        want_t, want_rms, want_ok, _ = ro.icp_point_to_plane(scan, ref, normals, init, d_max, sub, max_iter=max_iter,
                                                             rms_threshold=rms_threshold)
        assert np.allclose(t.rotation, want_t[0], atol=1e-9) and np.allclose(t.translation, want_t[1], atol=1e-9)
        assert np.isclose(rms, want_rms, rtol=1e-7) and bool(ok) == bool(want_ok)
    # the registration error after ICP is at the noise level of the pair
    true_rot = synthetic.rotation_from_rotvec(synthetic._PAIR_ROTVEC)
    assert np.abs(t.rotation - true_rot).max() < 1e-3
    assert torch.cuda.is_available()


def test_icp_point_to_point_against_oracle():
    """The reference's icp_point_to_point raises on every input (icp.py:118-120); the package provides
    what it intends, checked against the NumPy / KDTree restatement of the same loop."""
    from shot_fpfh_b200 import ops, synthetic
    from shot_fpfh_b200.core import RigidTransform
    from shot_fpfh_b200.device import upload
    from shot_fpfh_b200.icp import icp_point_to_point

    scan, ref, normals, scan_idx, ref_idx = registration_case()
    s = np.sqrt(3.0 / scan.shape[0])
    _, init = ro.ransac_on_matches(scan_idx, ref_idx, scan, ref, np.random.default_rng(seed=72), n_draws=300,
                                   distance_threshold=0.05)
    d_max = 4 * s
    sub = ops.voxel_subsample(upload(scan), 3 * s).cpu().numpy()
    for max_iter, rms_threshold in ((15, 1e-9), (60, 0.35)):  # runs out of iterations / stops early
        t, rms, ok = icp_point_to_point(scan, ref, RigidTransform(init[0], init[1]), d_max=d_max, voxel_size=3 * s,
                                        max_iter=max_iter, rms_threshold=rms_threshold)
        want_t, want_rms, want_ok, iterations = ro.icp_point_to_point(scan, ref, init, d_max, sub, max_iter=max_iter,
                                                                      rms_threshold=rms_threshold)
        assert np.allclose(t.rotation, want_t[0], atol=1e-9) and np.allclose(t.translation, want_t[1], atol=1e-9)
        assert np.isclose(rms, want_rms, rtol=1e-9) and bool(ok) == bool(want_ok)
    true_rot = synthetic.rotation_from_rotvec(synthetic._PAIR_ROTVEC)
    assert np.abs(t.rotation - true_rot).max() < 2e-3

"""
GPU parity of the "next" row `compute_normals` (pca_based_descriptors.py:29-59): k-nearest-neighbour sets against
sklearn's KDTree.query, normals against the oracle (NumPy eigh), including the LAPACK sign of normals that are not
re-oriented. Bar: kNN sets identical; normals within 1e-9 (float64 path) except near-degenerate neighbourhoods.
"""

import numpy as np
import pytest

from oracle import normals_oracle
from shot_fpfh_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_knn_sets_equal_kdtree():
    import torch

    from shot_fpfh_b200.descriptors.pca_based_descriptors import knn_device
    from shot_fpfh_b200.device import upload

    rng = np.random.default_rng(0)
    surface, _ = synthetic.bumpy_sphere(60000, seed=3)
    volume = rng.uniform(-1, 1, size=(40000, 3)) ** 3  # strongly non-uniform density
    for pts, k in ((surface, 30), (surface, 7), (volume, 30), (volume, 1)):
        q = np.concatenate([pts[::37], pts[:50] + 1e-3])
        want = normals_oracle.knn_sets(q, pts, k)
        got = knn_device(upload(pts), upload(q), k).cpu().numpy()
        assert got.shape == want.shape
        same = [set(a) == set(b) for a, b in zip(got, want)]
        assert all(same), f"{len(same) - sum(same)} of {len(same)} k-NN sets differ (k={k})"
        # nearest first
        d = np.linalg.norm(pts[got] - q[:, None, :], axis=2)
        assert np.all(np.diff(d, axis=1) >= 0)
    assert torch.cuda.is_available()


@pytest.mark.parametrize("with_pre", [True, False])
def test_compute_normals_knn_matches_oracle(with_pre):
    from shot_fpfh_b200.descriptors import compute_normals

    pts, dirs = synthetic.bumpy_sphere(30000, seed=9)
    pre = dirs if with_pre else None
    want = normals_oracle.compute_normals(pts[::5], pts, k=30, pre_computed_normals=pre[::5] if with_pre else None)
    got = compute_normals(pts[::5], pts, k=30, pre_computed_normals=pre[::5] if with_pre else None)
    assert got.shape == want.shape and got.dtype == np.float64
    diff = np.abs(got - want).max(axis=1)
    print(f"normals (k=30, pre={with_pre}): max abs diff {diff.max():.2e}, rows > 1e-9: {(diff > 1e-9).sum()}")
    assert (diff > 1e-9).sum() == 0
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0)


def test_compute_normals_radius_variant_and_c1_fixture():
    """The radius variant, and the C1 golden normals (reference compute_normals(k=30) rounded to float32)."""
    from conftest import load_golden

    from shot_fpfh_b200.descriptors import compute_normals

    pts, dirs = synthetic.bumpy_sphere(20000, seed=4)
    r = 4.0 * synthetic.mean_spacing(20000)
    want = normals_oracle.compute_normals(pts[::9], pts, radius=r, pre_computed_normals=dirs[::9])
    got = compute_normals(pts[::9], pts, radius=r, pre_computed_normals=dirs[::9])
    assert np.abs(got - want).max() < 1e-9
    g = load_golden("c1_pair_30k")
    scan, d = synthetic.bumpy_sphere(int(g["n_points"]), int(g["seed"]))
    got = compute_normals(scan, scan, k=30, pre_computed_normals=d)
    assert np.abs(got.astype(np.float32) - g["scan_normals_f32"]).max() < 1e-6

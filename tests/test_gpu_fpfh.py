"""
GPU parity, kernel group P: FPFH through `compute_fpfh_descriptor` (reference signature), against the golden
fixtures — 125-d from the unmodified reference, 33-d from the reference with the one-token fix at fpfh.py:78
(the unpatched reference raises) — and against the oracle. Bar: <= 1e-4 relative L2 per row.
"""

import os

import numpy as np
import pytest
from conftest import edge_case_inputs, golden_pair_inputs, load_golden, rel_l2

from oracle import fpfh_oracle
from shot_fpfh_b200 import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _check(got, want, what, max_bad=0):
    assert got.shape == want.shape and got.dtype == np.float64
    err = rel_l2(got, want)
    bad = int((err > TOL).sum())
    print(f"{what}: rows {err.shape[0]}, median {np.median(err):.2e}, max {err.max():.2e}, above {TOL:g}: {bad}")
    assert bad <= max_bad, f"{what}: {bad} rows above {TOL}"


@pytest.mark.parametrize("name", ["small_pair_4k", "c1_pair_30k"])
def test_golden_fpfh(name):
    from shot_fpfh_b200.descriptors import compute_fpfh_descriptor

    g = load_golden(name)
    clouds, radius = golden_pair_inputs(g)
    for tag, (cloud, normals) in clouds.items():
        kp = g[f"{tag}_kp_grid"]
        f125 = compute_fpfh_descriptor(kp, cloud, normals, radius=radius, n_bins=5, verbose=False)
        _check(f125, g[f"{tag}_fpfh125_grid"], f"{name}/{tag}/125")
        f33 = compute_fpfh_descriptor(kp, cloud, normals, radius, 11, decorrelated=True, verbose=False)
        assert f33.shape == (kp.shape[0], 33)
        _check(f33, g[f"{tag}_fpfh33_grid"], f"{name}/{tag}/33 (patched reference)")


def test_edge_cases_duplicates_counted_in_divisor():
    """Duplicated points: distance-0 neighbours are excluded from the sums but counted in K (fpfh.py:79, :115)."""
    from shot_fpfh_b200.descriptors import compute_fpfh_descriptor

    g = load_golden("edge_cases")
    pts, nrm, _, radius = edge_case_inputs(g)
    kp = g["edge_fpfh_kp"]
    _check(compute_fpfh_descriptor(kp, pts, nrm, radius, 5, verbose=False), g["edge_fpfh125"], "edge/125")
    _check(compute_fpfh_descriptor(kp, pts, nrm, radius, 11, True, verbose=False), g["edge_fpfh33"], "edge/33")
    with pytest.raises(IndexError):
        compute_fpfh_descriptor(np.array([pts.shape[0]]), pts, nrm, radius, 5, verbose=False)
    empty = compute_fpfh_descriptor(np.zeros(0, dtype=np.int64), pts, nrm, radius, 5, verbose=False)
    assert empty.shape == (0, 125)


@pytest.mark.parametrize("n_bins,decorrelated", [(11, True), (5, False), (3, False), (8, True), (11, False)])
def test_oracle_parity_all_points_are_queries(n_bins, decorrelated):
    """Every point a query (the C3 shape) on a 12k cloud with perturbed normals; all rows checked."""
    from shot_fpfh_b200.descriptors import compute_fpfh_descriptor

    n = 12000
    pts, dirs = synthetic.bumpy_sphere(n, seed=31)
    rng = np.random.default_rng(6)
    normals = dirs + 0.15 * rng.normal(size=dirs.shape)
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    radius = 5.0 * synthetic.mean_spacing(n)
    kp = np.arange(n)
    want = fpfh_oracle.fpfh(kp, pts, normals, radius, n_bins, decorrelated)
    got = compute_fpfh_descriptor(kp, pts, normals, radius, n_bins, decorrelated, verbose=False)
    _check(got, want, f"oracle/{n_bins}/{'dec' if decorrelated else 'cor'}")


@pytest.mark.parametrize("n_bins,decorrelated", [(11, True), (4, False), (10, True)])
def test_filtered_bins_equal_float64_bins_noisy_normals_and_flat_cloud(n_bins, decorrelated):
    """SPFH with the float32 filter == SPFH with every pair through float64, bit for bit: a 300k cloud with noisy
    (and non-unit) normals, and a flat lattice whose alpha is 0 up to rounding (an edge when n_bins is even)."""
    import torch
    from shot_fpfh_b200 import ops
    from shot_fpfh_b200.device import Grid, upload

    rng = np.random.default_rng(12)
    n = 300_000
    pts, dirs = synthetic.bumpy_sphere(n, seed=8)
    normals = (dirs + 0.3 * rng.normal(size=dirs.shape)) * rng.uniform(0.5, 2.0, size=(n, 1))
    gx, gy = np.meshgrid(np.arange(300.0), np.arange(300.0))
    flat = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1) * 0.01
    flat_n = np.tile([0.0, 0.0, 1.0], (flat.shape[0], 1))
    for cloud, nrm, radius in ((pts, normals, 5.0 * synthetic.mean_spacing(n)), (flat, flat_n, 0.035)):
        grid = Grid().build(upload(cloud), upload(nrm), radius)
        offsets, nbr, _, _ = ops.radius_csr(grid, None, radius)
        fast = ops.spfh(grid, offsets, nbr, n_bins, decorrelated)
        os.environ["SF_SPFH_EXACT"] = "1"
        try:
            exact = ops.spfh(grid, offsets, nbr, n_bins, decorrelated)
        finally:
            del os.environ["SF_SPFH_EXACT"]
        assert torch.equal(fast, exact)
        assert float(fast.sum()) > 0
        # the warp-per-tile kernel (flattened pairs) == the warp-per-point kernel, also on a sub-range of the cloud
        os.environ["SF_SPFH_NO_TILES"] = "1"
        try:
            per_point = ops.spfh(grid, offsets, nbr, n_bins, decorrelated)
        finally:
            del os.environ["SF_SPFH_NO_TILES"]
        assert torch.equal(fast, per_point)
        lo, cnt = 1000, 70_001
        sub_offsets = (offsets[lo : lo + cnt + 1] - offsets[lo]).contiguous()
        sub_nbr = nbr[int(offsets[lo]) : int(offsets[lo + cnt])].contiguous()
        assert torch.equal(ops.spfh(grid, sub_offsets, sub_nbr, n_bins, decorrelated, self_range=(lo, cnt)), fast[lo : lo + cnt])
        grid.close()


def test_widened_download_equals_a_float64_copy():
    """device.download_widened: float32 rows -> float64 host array == rows.double().cpu(), any shape / size."""
    import torch
    from shot_fpfh_b200.device import download_widened

    for shape in ((0, 33), (1, 33), (777, 125), (300_001, 33), (5_000_000,)):
        rows = torch.randn(shape, device="cuda", dtype=torch.float32)
        got = download_widened(rows)
        assert got.dtype == np.float64 and got.shape == tuple(shape)
        assert np.array_equal(got, rows.double().cpu().numpy())


def test_large_size_properties_1m():
    """C3 size: 1M points, every point a query, 33-d. Properties + a spot check against the oracle's formulas."""
    import torch
    from shot_fpfh_b200 import ops
    from shot_fpfh_b200.descriptors.fpfh import fpfh_device
    from shot_fpfh_b200.device import Grid, upload
    from sklearn.neighbors import KDTree

    n = 1_000_000
    pts, normals = synthetic.bumpy_sphere(n, seed=0)
    radius = 5.0 * synthetic.mean_spacing(n)
    p_dev, n_dev = upload(pts), upload(normals)
    grid = Grid().build(p_dev, n_dev, radius)
    kp = torch.arange(n, dtype=torch.int64, device=p_dev.device)
    out_dev, mean_k = fpfh_device(grid, kp, radius, 11, True, out_dtype=torch.float32)  # the fused driver
    out = out_dev.cpu().numpy().astype(np.float64)
    assert out.shape == (n, 33) and np.isfinite(out).all() and (out >= 0).all()
    assert 60 < mean_k < 90
    # SPFH rows: each of the three 11-bin blocks sums to (#binned)/K <= (K-1)/K
    offsets, nbr, _, dist = ops.radius_csr(grid, None, radius, want_dist=True)
    spfh_dev = ops.spfh(grid, offsets, nbr, 11, True)
    # the piecewise entry points (exact CSR, float64 distances) give the fused driver's rows: bit for bit when the
    # driver gathers per neighbour like they do, to float32 summation order with its float4-row kernel
    piecewise = ops.fpfh(grid, offsets, nbr, dist, spfh_dev, kp, out_dtype=torch.float32)
    assert torch.allclose(piecewise, out_dev, rtol=2e-5, atol=1e-7)
    os.environ["SF_FPFH_NO_ROWS4"] = "1"
    try:
        per_keypoint, _ = fpfh_device(grid, kp, radius, 11, True, out_dtype=torch.float32)
    finally:
        del os.environ["SF_FPFH_NO_ROWS4"]
    assert torch.equal(piecewise, per_keypoint)
    # a permuted subset of the points: same rows
    sub = torch.randperm(n, device=kp.device)[: (3 * n) // 4]
    sub_rows, _ = fpfh_device(grid, sub, radius, 11, True, out_dtype=torch.float32)
    assert torch.equal(sub_rows, out_dev[sub])
    del piecewise, per_keypoint, sub_rows
    assert abs(mean_k - float(offsets[-1].item()) / n) < 1e-9
    # the float32-filtered bins (sf_math.cuh::fpfh_bins_fast) == the float64 bins on all 74M pairs x 3 features
    os.environ["SF_SPFH_EXACT"] = "1"
    try:
        spfh_exact = ops.spfh(grid, offsets, nbr, 11, True)
        spfh125_exact = ops.spfh(grid, offsets, nbr, 5, False)
    finally:
        del os.environ["SF_SPFH_EXACT"]
    assert torch.equal(spfh_exact, spfh_dev)
    assert torch.equal(spfh125_exact, ops.spfh(grid, offsets, nbr, 5, False))
    del spfh_exact, spfh125_exact
    spfh = spfh_dev.cpu().numpy()
    blocks = spfh.reshape(n, 3, 11).sum(axis=2)
    assert (blocks <= 1.0 + 1e-6).all() and (blocks[:, 1] > 0.9).all()  # phi is always in range
    # spot check: 200 rows against the oracle evaluated on their two-ring neighbourhoods only
    rows = np.random.default_rng(1).choice(n, 200, replace=False)
    tree = KDTree(pts)
    edges = fpfh_oracle.bin_edges(11)
    ring1, dist1 = tree.query_radius(pts[rows], radius, return_distance=True)
    worst = 0.0
    for r, i in enumerate(rows):
        def spfh_of(j):
            nb = tree.query_radius(pts[j : j + 1], radius)[0]
            a, p, t = fpfh_oracle.pair_features(pts[j], normals[j], pts[nb], normals[nb])
            return fpfh_oracle.spfh_row(a, p, t, nb.shape[0], 11, True, edges)
        far = dist1[r] > 0
        acc = sum(spfh_of(j) / d for j, d in zip(ring1[r][far], dist1[r][far]))
        want = spfh_of(i) + acc / ring1[r].shape[0]
        worst = max(worst, float(rel_l2(out[i], want)))
    print(f"1M FPFH-33 spot check: worst rel-L2 {worst:.2e}")
    assert worst < TOL
    grid.close()

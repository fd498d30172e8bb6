"""
GPU parity, kernel group G: the uniform-grid radius search against sklearn's KDTree.query_radius (what the
reference calls) and against the golden neighbour lists. Bar: neighbour SETS bit-exact, distances bit-exact.
"""

import numpy as np
import pytest
from conftest import edge_case_inputs, golden_pair_inputs, load_golden

from oracle import neighbors_oracle
from shot_fpfh_b200 import synthetic

pytestmark = pytest.mark.gpu


def _csr_sets(search, queries, radius=None, return_distance=False):
    out = search.query_radius(queries, radius, return_distance=return_distance)
    return out


def _assert_same_sets(got, want):
    assert len(got) == len(want)
    for i in range(len(got)):
        assert np.array_equal(np.sort(got[i]), np.sort(want[i])), f"query {i}: neighbour sets differ"


def test_golden_neighbour_lists():
    from shot_fpfh_b200.neighbors import RadiusSearch

    g = load_golden("small_pair_4k")
    clouds, radius = golden_pair_inputs(g)
    for tag, (cloud, _) in clouds.items():
        kp = g[f"{tag}_kp_grid"]
        s = RadiusSearch(cloud, radius)
        got = s.query_radius(cloud[kp])
        offs, idx = g[f"{tag}_nbr_offsets"], g[f"{tag}_nbr_indices"]
        assert np.array_equal(np.cumsum([0] + [len(n) for n in got]), offs)
        for i in range(len(got)):
            assert np.array_equal(np.sort(got[i]), idx[offs[i] : offs[i + 1]])
        s.close()


def test_lattice_inclusive_boundary():
    """Distances exactly equal to the radius are inside (<=), as in sklearn."""
    from shot_fpfh_b200.neighbors import RadiusSearch

    g = load_golden("edge_cases")
    pts, q = g["lattice_points"], g["lattice_queries"]
    for name in ("half", "quarter", "diag"):
        r = float(g[f"lattice_{name}_r"])
        s = RadiusSearch(pts, r)
        got = s.query_radius(q)
        offs, idx = g[f"lattice_{name}_offsets"], g[f"lattice_{name}_indices"]
        for i in range(len(got)):
            assert np.array_equal(np.sort(got[i]), idx[offs[i] : offs[i + 1]]), (name, i)
        s.close()


def test_edge_queries_off_cloud_empty_and_duplicates():
    from shot_fpfh_b200.neighbors import RadiusSearch

    g = load_golden("edge_cases")
    pts, _, queries, radius = edge_case_inputs(g)
    s = RadiusSearch(pts, radius)
    got, dist = s.query_radius(queries, return_distance=True)
    offs, idx = g["edge_nbr_offsets"], g["edge_nbr_indices"]
    for i in range(len(got)):
        assert np.array_equal(np.sort(got[i]), idx[offs[i] : offs[i + 1]]), i
    assert len(got[90]) == 0 and len(got[91]) == 0  # the two far-away queries
    # distances: sqrt of the sequential float64 reduced distance, bit for bit
    for i in (0, 7, 65, 95):
        d = pts[got[i]] - queries[i]
        sq = d * d
        assert np.array_equal(dist[i], np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2]))
    s.close()


@pytest.mark.parametrize("n,seed", [(30000, 0), (200000, 4)])
def test_against_kdtree_all_points(n, seed):
    """Every cloud point as a query (the FPFH case) against the KD-tree, sets and distances."""
    from shot_fpfh_b200.neighbors import RadiusSearch

    pts, _ = synthetic.bumpy_sphere(n, seed)
    radius = 5.0 * synthetic.mean_spacing(n)
    stride = max(1, n // 20000)
    q = pts[::stride]
    want, wdist = neighbors_oracle.kdtree_radius(pts, q, radius, return_distance=True)
    s = RadiusSearch(pts, radius)
    got, gdist = s.query_radius(q, return_distance=True)
    _assert_same_sets(got, want)
    for i in range(0, len(got), 97):
        assert np.array_equal(gdist[i][np.argsort(got[i])], wdist[i][np.argsort(want[i])])
    # self-CSR (queries=None): counts agree with the explicit query path
    offsets, nbr_sorted, _, _ = s.csr(None, want_index=False, want_sorted=True)
    assert int(offsets[-1]) == sum(len(x) for x in s.query_radius(pts))
    s.close()


def test_smaller_radius_on_same_grid_and_random_box_cloud():
    """A volumetric (non-surface) cloud, queries outside the bounding box, radius below the build radius."""
    from shot_fpfh_b200.neighbors import RadiusSearch

    rng = np.random.default_rng(12)
    pts = rng.uniform(-1, 1, size=(50000, 3)) * np.array([1.0, 0.5, 2.0])
    q = np.concatenate([rng.uniform(-1.2, 1.2, size=(3000, 3)) * np.array([1.0, 0.5, 2.0]), pts[:500]])
    s = RadiusSearch(pts, 0.08)
    for r in (0.08, 0.05, 0.011):
        want = neighbors_oracle.kdtree_radius(pts, q, r)
        _assert_same_sets(s.query_radius(q, r), want)
    with pytest.raises(Exception):
        s.query_radius(q, 0.2)  # larger than the cell edge the grid was built for: refused, not silently wrong
    s.close()


def test_voxel_subsampling_equals_host_semantics():
    """G3: the device support reducer against the host restatement of grid_subsampling (core/subsampling.py:5-39)."""
    from shot_fpfh_b200.subsampling import grid_subsampling, grid_subsampling_gpu

    rng = np.random.default_rng(3)
    clouds = [
        synthetic.bumpy_sphere(200_000, seed=2)[0],
        rng.uniform(-3, 5, size=(50_000, 3)),
        np.round(rng.uniform(0, 1, size=(20_000, 3)), 2),  # many points exactly on voxel boundaries / duplicates
    ]
    for pts, voxels in zip(clouds, ((0.01, 0.05, 0.4), (0.3, 2.5), (0.05, 0.1, 0.25))):
        for voxel in voxels:
            want = grid_subsampling(pts, voxel)
            got = grid_subsampling_gpu(pts, voxel)
            assert got.shape == want.shape, (voxel, got.shape, want.shape)
            keys = ((pts - pts.min(axis=0)) // voxel).astype(int)
            assert np.array_equal(keys[got], keys[want])  # same voxels, same (lexicographic) order
            differ = np.nonzero(got != want)[0]
            # picks may only differ on exact distance ties (host and device sum the barycentre in the same order,
            # so in practice they do not differ at all)
            for i in differ[:50]:
                members = np.nonzero((keys == keys[got[i]]).all(axis=1))[0]
                centre = pts[members].mean(axis=0)
                assert abs(np.linalg.norm(pts[got[i]] - centre) - np.linalg.norm(pts[want[i]] - centre)) < 1e-12
            assert differ.shape[0] <= 0.01 * got.shape[0]
    assert grid_subsampling_gpu(np.zeros((0, 3)), 0.1).shape == (0,)


def test_keypoint_selectors_against_oracle():
    """§8f row 1: select_keypoints_subsampling / select_keypoints_with_density_threshold (keypoint_selection.py:34-122)."""
    from oracle import keypoints_oracle
    from shot_fpfh_b200.keypoint_selection import select_keypoints_subsampling, select_keypoints_with_density_threshold

    pts = synthetic.bumpy_sphere(30_000, seed=4)[0]
    s = synthetic.mean_spacing(30_000)

    def same_up_to_ties(got, want, keys):
        """Same voxels in the same order; representatives may differ only where both are equidistant from the voxel's
        barycentre (2-point voxels: the reference's pick then follows an unstable argsort and summation order)."""
        got, want = got.astype(int), want.astype(int)
        assert got.shape == want.shape and np.array_equal(keys[got], keys[want])
        for i in np.nonzero(got != want)[0]:
            members = np.nonzero((keys == keys[got[i]]).all(axis=1))[0]
            centre = pts[members].mean(axis=0)
            assert abs(np.linalg.norm(pts[got[i]] - centre) - np.linalg.norm(pts[want[i]] - centre)) < 1e-12
        assert (got != want).mean() < 0.05

    for voxel in (2.0 * s, 3.5 * s):
        keys = ((pts - pts.min(axis=0)) // voxel).astype(int)
        got_sub = select_keypoints_subsampling(pts, voxel)
        want_sub = keypoints_oracle.select_keypoints_subsampling(pts, voxel)
        same_up_to_ties(got_sub, want_sub, keys)
        ambiguous = {tuple(k) for k in keys[got_sub[got_sub != want_sub]]}  # voxels whose representative is a tie

        def settled(idx):
            return np.array([i for i in idx.astype(int) if tuple(keys[i]) not in ambiguous])

        for value, radius in ((3, None), (9, voxel), (25, 1.5 * voxel), (6, 0.6 * voxel), (10**6, None)):
            want = keypoints_oracle.select_keypoints_with_density_threshold(pts, voxel, value, radius)
            got = select_keypoints_with_density_threshold(pts, voxel, value, radius)
            if radius is None or radius == voxel:  # thresholds the voxel's own population: independent of the pick
                assert got.shape == want.shape, (voxel, value, radius, got.shape, want.shape)
                if want.shape[0]:
                    same_up_to_ties(got, want, keys)
            else:  # thresholds the count AROUND the representative: compare where the representative is settled
                assert np.array_equal(settled(got), settled(want)) and settled(want).shape[0] > 1000
    assert select_keypoints_with_density_threshold(np.zeros((0, 3)), 0.1, 3).shape == (0,)


def test_counting_sort_build_equals_radix_sort_build():
    """The grid's counting sort + stable rank gives the permutation of the stable radix sort (ascending original index
    inside a cell): random surface, duplicated points, one crowded cell, a large radius (few, crowded cells)."""
    import os

    import torch
    from shot_fpfh_b200 import ops
    from shot_fpfh_b200.device import Grid, upload

    rng = np.random.default_rng(3)
    pts, _ = synthetic.bumpy_sphere(200_000, seed=4)
    dup = pts.copy()
    dup[1000:3000] = dup[17]           # 2000 copies of one point
    dup[5000:9000] = dup[5000:9000].round(2)  # many exact duplicates on a lattice
    cases = [(pts, 5.0 * synthetic.mean_spacing(200_000)), (dup, 5.0 * synthetic.mean_spacing(200_000)),
             (rng.random((30_000, 3)), 0.4), (np.zeros((5000, 3)), 1.0)]
    for cloud, radius in cases:
        p = upload(cloud)
        perms = []
        for radix in ("0", "1"):
            os.environ["SF_GRID_RADIX"] = radix
            try:
                grid = Grid().build(p, p, radius)
                perm, inv = ops.grid_permutation(grid)
                offsets, nbr, _, _ = ops.radius_csr(grid, p[:2000].contiguous(), radius)
                perms.append((perm.clone(), inv.clone(), offsets.clone(), nbr.clone()))
                grid.close()
            finally:
                del os.environ["SF_GRID_RADIX"]
        for x, y in zip(*perms):
            assert torch.equal(x, y)
        assert torch.equal(torch.sort(perms[0][0]).values, torch.arange(cloud.shape[0], device=p.device, dtype=torch.int32))


def test_degenerate_clouds():
    from shot_fpfh_b200.neighbors import RadiusSearch

    one = np.array([[0.5, -1.0, 2.0]])
    s = RadiusSearch(one, 0.1)
    got = s.query_radius(np.array([[0.5, -1.0, 2.0], [0.5, -1.0, 2.2]]))
    assert np.array_equal(got[0], [0]) and len(got[1]) == 0
    s.close()
    flat = np.zeros((1000, 3))
    flat[:, 0] = np.linspace(0, 1, 1000)  # all points on a line: two grid dimensions collapse to one cell
    s = RadiusSearch(flat, 0.01)
    want = neighbors_oracle.kdtree_radius(flat, flat[::10], 0.01)
    _assert_same_sets(s.query_radius(flat[::10]), want)
    assert len(s.query_radius(np.zeros((0, 3)))) == 0
    s.close()

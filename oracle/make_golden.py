"""
TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (imported in
place from /root/reference, build container only) on seeded synthetic inputs, and pins the oracle
restatement against it in the same run (bit-exact where the arithmetic is restated op for op).

    python -m oracle.make_golden            # ~2 min on 8 cores

Inputs are regenerated from seeds by `shot_fpfh_b200.synthetic`; only what cannot be regenerated without
the reference (PCA normals, barycentre-closest keypoints) and the reference's OUTPUTS are stored.
"""

from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fpfh_oracle, matching_oracle, neighbors_oracle, shot_oracle  # noqa: E402
from oracle.reference_harness import import_reference, patched_fpfh_decorrelated  # noqa: E402
from shot_fpfh_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
MIN_NB = 10
report: dict = {"generated_with": {}, "checks": {}}


def rel_l2(a, b):
    num = np.linalg.norm(a - b, axis=-1)
    den = np.maximum(np.linalg.norm(b, axis=-1), 1e-300)
    return num / den


def ref_shot(ref, cloud, normals, keypoints, radius, normalize=True, min_nb=MIN_NB, n_procs=8):
    from shot_fpfh.descriptors import ShotMultiprocessor

    with ShotMultiprocessor(
        normalize=normalize, min_neighborhood_size=min_nb, n_procs=n_procs, disable_progress_bar=True, verbose=False
    ) as mp:
        return mp.compute_descriptor_single_scale(cloud, normals, keypoints, radius)


def ref_lrfs(cloud, keypoints, radius):
    from shot_fpfh.descriptors.shot import get_local_rf
    from sklearn.neighbors import KDTree

    nbh = KDTree(cloud).query_radius(keypoints, radius)
    return np.array([get_local_rf((kp, cloud[nbh[i]], radius)) for i, kp in enumerate(keypoints)])


def build_pair(ref, n_points, seed, pca_normals: bool):
    from shot_fpfh.descriptors import compute_normals

    scan, dirs = synthetic.bumpy_sphere(n_points, seed)
    if pca_normals:
        # reference compute_normals (pca_based_descriptors.py:29-59), k = 30 as the CLI does; stored rounded to
        # float32 and widened again so that the fixture stays small AND is the exact input of the reference.
        normals = compute_normals(scan, scan, k=30, pre_computed_normals=dirs)
        normals = normals.astype(np.float32).astype(np.float64)
    else:
        normals = dirs
    ref_pts, ref_normals, perm, rot, trans = synthetic.rigid_pair(scan, normals)
    return scan, normals, ref_pts, ref_normals, perm


def golden_pair(ref, name, n_points, dense_stride, fpfh_all_points: bool, store_dense: bool):
    from shot_fpfh.core import grid_subsampling
    from shot_fpfh.descriptors import compute_fpfh_descriptor
    from shot_fpfh.matching import basic_matching, match_descriptors, threshold_filter

    t0 = time.time()
    fpfh33 = patched_fpfh_decorrelated()
    scan, normals, ref_pts, ref_normals, perm = build_pair(ref, n_points, 0, pca_normals=True)
    radius = 5.0 * synthetic.mean_spacing(n_points)
    out = {
        "n_points": n_points,
        "seed": 0,
        "radius": radius,
        "min_neighborhood_size": MIN_NB,
        "scan_normals_f32": normals.astype(np.float32),
    }
    clouds = {"scan": (scan, normals), "ref": (ref_pts, ref_normals)}
    desc = {}
    for tag, (cloud, nrm) in clouds.items():
        kp_grid = np.asarray(grid_subsampling(cloud, 2.0 * radius), dtype=np.int64)
        kp_dense = np.arange(0, n_points, dense_stride, dtype=np.int64)
        out[f"{tag}_kp_grid"] = kp_grid
        for kp_name, kp in (("grid", kp_grid), ("dense", kp_dense)):
            d = ref_shot(ref, cloud, nrm, cloud[kp], radius)
            desc[(tag, kp_name)] = d
            if kp_name == "grid" or store_dense:
                out[f"{tag}_shot_{kp_name}"] = d
            else:  # too large to commit: the tests rebuild it with the (bit-exact) oracle and check this digest
                out[f"{tag}_shot_{kp_name}_sha256"] = np.frombuffer(
                    hashlib.sha256(np.ascontiguousarray(d).tobytes()).digest(), dtype=np.uint8
                )
            # pin the oracle: bit-exact descriptor and LRF
            od, olrf = shot_oracle.shot_single_scale(cloud, nrm, cloud[kp], radius, True, MIN_NB, return_lrf=True)
            rl = ref_lrfs(cloud, cloud[kp], radius)
            report["checks"][f"{name}/{tag}/shot_{kp_name}/oracle_max_abs_diff"] = float(np.abs(od - d).max())
            report["checks"][f"{name}/{tag}/lrf_{kp_name}/oracle_max_abs_diff"] = float(np.abs(olrf - rl).max())
            assert np.array_equal(od, d), "SHOT oracle is not bit-exact against the reference"
            assert np.array_equal(olrf, rl), "LRF oracle is not bit-exact against the reference"
            if kp_name == "grid":
                out[f"{tag}_lrf_grid"] = rl
        # un-normalised variant and the all-zero default (F4) on the grid keypoints
        out[f"{tag}_shot_grid_raw"] = ref_shot(ref, cloud, nrm, cloud[kp_grid], radius, normalize=False)
        assert not ref_shot(ref, cloud, nrm, cloud[kp_grid][:40], radius, min_nb=100).any()
        # neighbour lists of the grid keypoints
        nbh = neighbors_oracle.kdtree_radius(cloud, cloud[kp_grid], radius)
        offs, idx, _ = neighbors_oracle.to_sorted_csr(nbh)
        out[f"{tag}_nbr_offsets"], out[f"{tag}_nbr_indices"] = offs, idx.astype(np.int32)
        # FPFH: 125-d from the unmodified reference, 33-d from the patched one
        kp_f = np.arange(n_points, dtype=np.int64) if fpfh_all_points else kp_grid
        f125 = compute_fpfh_descriptor(kp_f, cloud, nrm, radius=radius, n_bins=5, verbose=False)
        f33 = fpfh33(kp_f, cloud, nrm, radius=radius, n_bins=11, decorrelated=True, verbose=False)
        o125 = fpfh_oracle.fpfh(kp_f, cloud, nrm, radius, 5, False)
        o33 = fpfh_oracle.fpfh(kp_f, cloud, nrm, radius, 11, True)
        report["checks"][f"{name}/{tag}/fpfh125/oracle_max_rel_l2"] = float(rel_l2(o125, f125).max())
        report["checks"][f"{name}/{tag}/fpfh33/oracle_max_rel_l2"] = float(rel_l2(o33, f33).max())
        assert rel_l2(o125, f125).max() < 1e-12 and rel_l2(o33, f33).max() < 1e-12
        sel = kp_grid if fpfh_all_points else np.arange(kp_grid.shape[0])
        out[f"{tag}_fpfh125_grid"] = f125[sel]
        out[f"{tag}_fpfh33_grid"] = f33[sel]
        try:
            compute_fpfh_descriptor(kp_grid[:3], cloud, nrm, radius=radius, n_bins=11, decorrelated=True, verbose=False)
            raise AssertionError("the unpatched reference was expected to raise on decorrelated=True (F2)")
        except ValueError:
            report["checks"][f"{name}/{tag}/fpfh33_unpatched_raises"] = True

    # matching (dense SHOT rows, with a few rows zeroed to exercise the non-empty filter)
    a, b = desc[("scan", "dense")].copy(), desc[("ref", "dense")].copy()
    a[::17] = 0.0
    b[5::23] = 0.0
    out["dense_stride"] = dense_stride
    out["match_zeroed_scan_stride"], out["match_zeroed_ref_stride"] = 17, np.array([5, 23])
    m = basic_matching(a, b)
    out["basic_scan"], out["basic_ref"] = m[0].astype(np.int64), m[1].astype(np.int64)
    om = matching_oracle.basic_matching(a, b)
    assert np.array_equal(om[0], m[0]) and np.array_equal(om[1], m[1])
    for mult in (1.5, 3.0):
        for recip in (False, True):
            m = match_descriptors(
                a, b, threshold_filter, filter_nonreciprocal=recip, verbose=False, n_min_matches=10,
                threshold_multiplier=mult,
            )
            key = f"thr{mult}_{'recip' if recip else 'all'}"
            out[f"{key}_scan"], out[f"{key}_ref"] = m[0].astype(np.int64), m[1].astype(np.int64)
            om = matching_oracle.match_descriptors(
                a, b, matching_oracle.threshold_filter, recip, 10, threshold_multiplier=mult
            )
            assert np.array_equal(om[0], m[0]) and np.array_equal(om[1], m[1])
    report["checks"][f"{name}/matching/oracle_exact"] = True
    # F3: the reference's ratio matcher raises
    from shot_fpfh.matching import double_matching_with_rejects

    try:
        double_matching_with_rejects(a, b, 0.8, verbose=False)
        report["checks"][f"{name}/ratio_matcher_raises"] = False
    except Exception as exc:  # noqa: BLE001
        report["checks"][f"{name}/ratio_matcher_raises"] = type(exc).__name__
    rs, rr = matching_oracle.ratio_matching(a, b, 0.8)
    out["ratio0.8_scan_RESTATEMENT"], out["ratio0.8_ref_RESTATEMENT"] = rs, rr

    # F1: the accumulating histogram is NOT the reference
    from sklearn.neighbors import KDTree

    kp = scan[out["scan_kp_grid"][:50]]
    nbh = KDTree(scan).query_radius(kp, radius)
    devs = []
    for i, p in enumerate(kp):
        pts, nr = scan[nbh[i]], normals[nbh[i]]
        lrf = shot_oracle.local_reference_frame(p, pts, radius)
        rho = np.linalg.norm(pts - p, axis=1)
        keep = rho > 0
        order = np.argsort(rho[keep])
        local = ((pts[keep] - p) @ lrf)[order]
        cosine = np.clip(nr[keep] @ lrf[:, 2], -1, 1)[order]
        acc = shot_oracle.apply_accumulate(shot_oracle.shot_writes(local, cosine, rho[keep][order], radius))
        acc /= np.linalg.norm(acc)
        devs.append(float(rel_l2(acc, out["scan_shot_grid"][i])))
    report["checks"][f"{name}/accumulate_vs_reference_median_rel_l2"] = float(np.median(devs))

    np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **out)
    print(f"{name}: done in {time.time() - t0:.1f}s")


def golden_edge_cases(ref):
    """Degenerate inputs the reference tolerates silently (SURVEY.md §8b error conventions)."""
    from shot_fpfh.descriptors import compute_fpfh_descriptor

    out = {}
    fpfh33 = patched_fpfh_decorrelated()
    # (1) lattice: distances exactly equal to the radius (inclusive predicate), many exact ties
    g = np.arange(6) * 0.25
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    q = lattice[[0, 43, 86, 129, 215]]
    for r_name, r in (("half", 0.5), ("quarter", 0.25), ("diag", float(np.sqrt(0.125)))):
        nbh = neighbors_oracle.kdtree_radius(lattice, q, r)
        offs, idx, _ = neighbors_oracle.to_sorted_csr(nbh)
        o2, i2, _ = neighbors_oracle.brute_force_radius(lattice, q, r)
        assert np.array_equal(offs, o2) and np.array_equal(idx, i2), "brute-force predicate != KDTree"
        out[f"lattice_{r_name}_r"], out[f"lattice_{r_name}_offsets"], out[f"lattice_{r_name}_indices"] = r, offs, idx
    out["lattice_points"], out["lattice_queries"] = lattice, q

    # (2) sparse / empty / duplicated neighbourhoods on a small bumpy sphere
    n = 3000
    pts, dirs = synthetic.bumpy_sphere(n, seed=7)
    pts = np.concatenate([pts, pts[:40]])  # exact duplicates: distance-0 neighbours that are not the query
    nrm = np.concatenate([dirs, dirs[:40]])
    radius = 5.0 * synthetic.mean_spacing(n)
    rng = np.random.default_rng(11)
    queries = np.concatenate(
        [
            pts[:60],  # includes the duplicated points
            pts[100:130] + rng.normal(scale=0.3 * radius, size=(30, 3)),  # off-surface queries (not in the cloud)
            np.array([[5.0, 5.0, 5.0], [-3.0, 0.0, 0.0]]),  # empty neighbourhoods
            pts[200:210] * 1.04,  # sparse neighbourhoods (below min_neighborhood_size for some)
        ]
    )
    out["edge_n"], out["edge_seed"], out["edge_radius"], out["edge_queries"] = n, 7, radius, queries
    for min_nb in (10, 40):
        d = ref_shot(ref, pts, nrm, queries, radius, min_nb=min_nb, n_procs=2)
        od = shot_oracle.shot_single_scale(pts, nrm, queries, radius, True, min_nb)
        assert np.array_equal(od, d)
        out[f"edge_shot_minnb{min_nb}"] = d
    out["edge_lrf"] = ref_lrfs(pts, queries, radius)
    nbh = neighbors_oracle.kdtree_radius(pts, queries, radius)
    out["edge_nbr_offsets"], idx, _ = neighbors_oracle.to_sorted_csr(nbh)
    out["edge_nbr_indices"] = idx.astype(np.int32)
    kp = np.concatenate([np.arange(0, 60), np.arange(n, n + 40)]).astype(np.int64)
    out["edge_fpfh_kp"] = kp
    out["edge_fpfh125"] = compute_fpfh_descriptor(kp, pts, nrm, radius=radius, n_bins=5, verbose=False)
    out["edge_fpfh33"] = fpfh33(kp, pts, nrm, radius=radius, n_bins=11, decorrelated=True, verbose=False)
    assert rel_l2(fpfh_oracle.fpfh(kp, pts, nrm, radius, 5, False), out["edge_fpfh125"]).max() < 1e-12
    assert rel_l2(fpfh_oracle.fpfh(kp, pts, nrm, radius, 11, True), out["edge_fpfh33"]).max() < 1e-12
    # serial debug twin (shot.py:310-499)
    from shot_fpfh.descriptors.shot import compute_shot_descriptor

    ds = compute_shot_descriptor(queries, pts, nrm, radius, min_neighborhood_size=10)
    assert np.array_equal(shot_oracle.shot_serial_debug(queries, pts, nrm, radius, 10), ds)
    out["edge_shot_serial"] = ds

    # (3) azimuth / cosine-bin boundary table (F6)
    from shot_fpfh.descriptors.shot import get_azimuth_idx

    s = np.sqrt(0.5)
    bx = np.array([0.0, -1.0, -s, 0.0, s, 1.0, s, 0.0, -s, 0.3, -0.2, 0.7, -0.9])
    by = np.array([0.0, 0.0, -s, -1.0, -s, 0.0, s, 1.0, s, 0.4, 0.5, -0.1, -0.8])
    out["azimuth_x"], out["azimuth_y"] = bx, by
    out["azimuth_idx"] = np.asarray(get_azimuth_idx(bx, by), dtype=np.int64)
    assert np.array_equal(shot_oracle.azimuth_octant(bx, by), out["azimuth_idx"])
    np.savez_compressed(os.path.join(GOLDEN, "edge_cases.npz"), **out)
    print("edge_cases: done")


def main():
    import scipy
    import sklearn

    os.makedirs(GOLDEN, exist_ok=True)
    ref = import_reference()
    report["generated_with"] = {
        "reference": "aubin-tchoi/shot-fpfh 1.1.0 (unmodified, /root/reference)",
        "numpy": np.__version__,
        "scipy": scipy.__version__,
        "scikit_learn": sklearn.__version__,
        "pinned_by_reference": {"numpy": "1.26.4", "scipy": "1.14.0", "scikit_learn": "1.5.1"},
    }
    golden_edge_cases(ref)
    golden_pair(ref, "small_pair_4k", 4000, 8, fpfh_all_points=True, store_dense=True)
    golden_pair(ref, "c1_pair_30k", 30000, 6, fpfh_all_points=False, store_dense=False)
    with open(os.path.join(GOLDEN, "PINNING_REPORT.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(report["checks"], indent=1, sort_keys=True))


if __name__ == "__main__":
    main()

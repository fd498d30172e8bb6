"""
The reference's registration flow (scripts/register_point_clouds.py:80-127 -> pipeline.py: select_keypoints,
compute_descriptors, find_descriptors_matches, run_ransac, run_icp) through THIS package's reference-shaped API, host
NumPy arrays in and out at every stage, on a synthetic rigid pair (SURVEY.md 8d generator). BASELINE.json configs[4]
names a 10M-point pair; the size is an argument:

    python scripts/run_registration_flow.py [n_points] [fpfh]

Prints wall-clock per stage and the error of the recovered transform against the known one. (`/root/reference` is not
on the GPU box, so its `RegistrationPipeline` object cannot be run there; `dropin.install()` binds exactly these
callables into it.)
"""

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from shot_fpfh_b200 import synthetic  # noqa: E402
from shot_fpfh_b200.core import RigidTransform  # noqa: E402
from shot_fpfh_b200.descriptors import ShotMultiprocessor, compute_fpfh_descriptor  # noqa: E402
from shot_fpfh_b200.icp import icp_point_to_plane  # noqa: E402
from shot_fpfh_b200.keypoint_selection import select_keypoints_subsampling  # noqa: E402
from shot_fpfh_b200.matching import match_descriptors, ransac_on_matches  # noqa: E402


def run(n: int, with_fpfh: bool = False, verbose: bool = True) -> dict:
    times = {}

    def stage(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        times[name] = time.perf_counter() - t0
        if verbose:
            print(f"  {name:34s} {1e3 * times[name]:10.1f} ms")
        return out

    scan, directions = synthetic.bumpy_sphere(n, seed=0)
    normals = synthetic.bumpy_sphere_true_normals(directions)  # the surface's own normals: ICP needs them (see there)
    ref, ref_normals, perm, rot, trans = synthetic.rigid_pair(scan, normals)
    s = synthetic.mean_spacing(n)
    radius = 5.0 * s
    if verbose:
        print(f"pair of {n} points, radius {radius:.5f}")
    scan_kp = stage("select_keypoints (scan)", lambda: select_keypoints_subsampling(scan, 3.75 * s))
    ref_kp = stage("select_keypoints (ref)", lambda: select_keypoints_subsampling(ref, 3.75 * s))
    with ShotMultiprocessor(min_neighborhood_size=10, verbose=False) as shot:
        scan_desc = stage("SHOT single scale (scan)", lambda: shot.compute_descriptor_single_scale(
            scan, normals, scan[scan_kp], radius))
        ref_desc = stage("SHOT single scale (ref)", lambda: shot.compute_descriptor_single_scale(
            ref, ref_normals, ref[ref_kp], radius))
    if with_fpfh:
        stage("FPFH 33-d on the keypoints (scan)", lambda: compute_fpfh_descriptor(
            scan_kp, scan, normals, radius, n_bins=11, decorrelated=True, verbose=False))
    # the two clouds' keypoints are different points of the surface (voxel selection of a rotated copy): reciprocity is
    # the filter that fits, as in pipeline.py:376-399 with `filter_nonreciprocal`
    matches = stage("match_descriptors (reciprocal)", lambda: match_descriptors(
        scan_desc, ref_desc, None, filter_nonreciprocal=True, verbose=False))
    del scan_desc, ref_desc
    import shot_fpfh_b200.matching.ransac as ransac_module

    ransac_module.rng = np.random.default_rng(seed=72)  # the reference's module-level generator (ransac.py:14)
    ratio, coarse = stage("ransac_on_matches (10000 draws)", lambda: ransac_on_matches(
        matches[0], matches[1], scan[scan_kp], ref[ref_kp], n_draws=10_000, distance_threshold=4 * s))
    fine, rms, _ = stage("icp_point_to_plane (<= 20 iters)", lambda: icp_point_to_plane(
        scan, ref, ref_normals, RigidTransform(coarse.rotation, coarse.translation), d_max=2 * s, voxel_size=3.75 * s,
        max_iter=20, rms_threshold=1e-9))
    out = {
        "n_points": n, "keypoints": (int(scan_kp.shape[0]), int(ref_kp.shape[0])), "matches": int(matches[0].shape[0]),
        "ransac_inlier_ratio": float(ratio),
        "coarse_rotation_error": float(np.abs(coarse.rotation - rot).max()),
        "fine_rotation_error": float(np.abs(fine.rotation - rot).max()),
        "fine_translation_error": float(np.abs(fine.translation - trans).max()),
        "seconds": times, "total_seconds": float(sum(times.values())),
        "peak_device_GiB": torch.cuda.max_memory_allocated() / 2**30,
    }
    if verbose:
        print({k: v for k, v in out.items() if k != "seconds"})
    return out


if __name__ == "__main__":
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    run(size, with_fpfh=len(sys.argv) > 2 and sys.argv[2] == "fpfh")

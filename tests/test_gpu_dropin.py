"""
GPU: the UNMODIFIED reference's own orchestration — `shot_fpfh.pipeline.RegistrationPipeline` from baseline/_ref (the
copy installed by baseline/install_reference.py) — run on top of this package through `dropin.install()`:
select_keypoints -> compute_descriptors -> find_descriptors_matches -> run_ransac -> run_icp, exactly the calls of
scripts/register_point_clouds.py:80-127 (BASELINE.json north_star: "register_point_clouds and pipeline.py run unchanged
on top of it"). Skipped only when baseline/_ref was not shipped.
"""

import importlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

sys.path.insert(0, ROOT)
from baseline import reference_loader  # noqa: E402

needs_reference = pytest.mark.skipif(not reference_loader.available(), reason="baseline/_ref was not shipped")


@pytest.fixture()
def reference_pipeline_on_b200():
    reference_loader.load()
    import shot_fpfh_b200.dropin as dropin

    done = dropin.install()
    assert len(done) >= 30
    yield importlib.import_module("shot_fpfh.pipeline")
    dropin.uninstall()


@needs_reference
def test_reference_pipeline_object_registers_a_pair(reference_pipeline_on_b200):
    from shot_fpfh_b200 import synthetic
    import shot_fpfh_b200.matching.ransac as ransac_module

    pipeline = reference_pipeline_on_b200
    ref_core = importlib.import_module("shot_fpfh.core")
    n = 60_000
    scan, directions = synthetic.bumpy_sphere(n, seed=0)
    normals = synthetic.bumpy_sphere_true_normals(directions)
    ref, ref_normals, perm, rot, trans = synthetic.rigid_pair(scan, normals)
    s = synthetic.mean_spacing(n)
    pipe = pipeline.RegistrationPipeline(scan=scan, scan_normals=normals, ref=ref, ref_normals=ref_normals)
    pipe.select_keypoints("subsampling", neighborhood_size=3.75 * s)
    assert pipe.scan_keypoints.shape[0] > 3000 and pipe.ref_keypoints.shape[0] > 3000
    pipe.compute_descriptors(radius=5.0 * s, descriptor_choice="shot_single_scale", subsample_support=False,
                             min_neighborhood_size=10, disable_progress_bars=True, verbose=False)
    assert pipe.scan_descriptors.shape == (pipe.scan_keypoints.shape[0], 352) and pipe.scan_descriptors.dtype == np.float64
    pipe.find_descriptors_matches("threshold", reject_threshold=0.8, threshold_multiplier=4.0)
    assert pipe.matches[0].shape[0] > 200 and pipe.matches[0].shape == pipe.matches[1].shape
    ransac_module.rng = np.random.default_rng(seed=72)
    coarse, ratio = pipe.run_ransac(n_draws=5000, max_inliers_distance=4 * s, disable_progress_bar=True)
    assert ratio > 0.05 and np.abs(coarse.rotation - rot).max() < 0.1
    # run_icp with the pipeline's own result, then with a REFERENCE RigidTransform (as_row only existed on
    # this package's class)
    fine, rms, _ = pipe.run_icp("point_to_plane", coarse, d_max=2 * s, voxel_size=3.75 * s, max_iter=20,
                                rms_threshold=1e-9, disable_progress_bar=True)
    assert np.abs(fine.rotation - rot).max() < 1e-5 and np.abs(fine.translation - trans).max() < 1e-5
    init = ref_core.RigidTransform(np.asarray(coarse.rotation), np.asarray(coarse.translation))
    fine2, _, _ = pipe.run_icp("point_to_plane", init, d_max=2 * s, voxel_size=3.75 * s, max_iter=20, rms_threshold=1e-9,
                               disable_progress_bar=True)
    assert np.abs(fine2.rotation - rot).max() < 1e-5
    # FPFH through the pipeline (the reference never passes `decorrelated`: the 125-d layout) and simple matching
    pipe.scan_descriptors = pipe.ref_descriptors = pipe.matches = None
    pipe.compute_descriptors(radius=5.0 * s, descriptor_choice="fpfh", fpfh_n_bins=5, disable_progress_bars=True,
                             verbose=False)
    assert pipe.scan_descriptors.shape == (pipe.scan_keypoints.shape[0], 125)
    pipe.find_descriptors_matches("simple", reject_threshold=0.8, threshold_multiplier=4.0)
    assert pipe.matches[0].shape[0] > 1000


@needs_reference
def test_rebound_functions_equal_the_reference_functions_on_the_same_inputs(reference_pipeline_on_b200):
    """The same calls through the reference's OWN functions (CPU) and through the rebound ones (GPU), small sizes."""
    import shot_fpfh_b200.dropin as dropin
    from conftest import rel_l2
    from shot_fpfh_b200 import synthetic

    n = 6000
    pts, directions = synthetic.bumpy_sphere(n, seed=4)
    normals = synthetic.bumpy_sphere_true_normals(directions)
    radius = 5.0 * synthetic.mean_spacing(n)
    kp = np.arange(0, n, 25)
    descriptors = importlib.import_module("shot_fpfh.descriptors")
    matching = importlib.import_module("shot_fpfh.matching")
    with descriptors.ShotMultiprocessor(min_neighborhood_size=10, verbose=False, disable_progress_bar=True) as shot:
        got = shot.compute_descriptor_single_scale(pts, normals, pts[kp], radius)
    got_f = descriptors.compute_fpfh_descriptor(kp, pts, normals, radius=radius, n_bins=5, disable_progress_bars=True,
                                                verbose=False)
    got_m = matching.basic_matching(got, got[::-1].copy())
    dropin.uninstall()  # the reference's own code again
    try:
        with descriptors.ShotMultiprocessor(min_neighborhood_size=10, verbose=False, disable_progress_bar=True,
                                            n_procs=4) as shot:
            want = shot.compute_descriptor_single_scale(pts, normals, pts[kp], radius)
        want_f = descriptors.compute_fpfh_descriptor(kp, pts, normals, radius=radius, n_bins=5,
                                                     disable_progress_bars=True, verbose=False)
        want_m = matching.basic_matching(got, got[::-1].copy())
    finally:
        dropin.install()
    keep = want.any(axis=1)
    assert np.array_equal(got.any(axis=1), keep)
    assert rel_l2(got[keep], want[keep]).max() < 1e-4
    assert rel_l2(got_f, want_f).max() < 1e-4
    assert np.array_equal(got_m[0], want_m[0]) and np.array_equal(got_m[1], want_m[1])

"""
GPU: the registration flow of the reference's pipeline (pipeline.py: select_keypoints -> compute_descriptors ->
find_descriptors_matches -> run_ransac -> run_icp) through this package's reference-shaped entry points, in the order
and with the argument conventions `RegistrationPipeline` uses (SHOT takes keypoint COORDINATES, pipeline.py:159; FPFH
takes INDICES, :330; matches index the descriptor rows, :428), on a rigid pair whose transform is known.
"""

import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_registration_flow_recovers_the_rigid_transform():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import run_registration_flow as flow

    out = flow.run(60_000, with_fpfh=True, verbose=False)
    print(out)
    assert out["keypoints"][0] > 3000 and out["matches"] > 200
    assert out["ransac_inlier_ratio"] > 0.1
    assert out["coarse_rotation_error"] < 0.1
    assert out["fine_rotation_error"] < 1e-5 and out["fine_translation_error"] < 1e-5


def test_fpfh_on_keypoint_subset_equals_rows_of_the_full_result():
    """pipeline.py:329-347 asks FPFH for the keypoints only: the rows equal those of the all-points call."""
    from shot_fpfh_b200 import synthetic
    from shot_fpfh_b200.descriptors import compute_fpfh_descriptor

    n = 50_000
    pts, normals = synthetic.bumpy_sphere(n, seed=3)
    radius = 5.0 * synthetic.mean_spacing(n)
    kp = np.random.default_rng(0).choice(n, 4321, replace=False)
    full = compute_fpfh_descriptor(np.arange(n), pts, normals, radius, 11, True, verbose=False)
    part = compute_fpfh_descriptor(kp, pts, normals, radius, 11, True, verbose=False)
    assert np.array_equal(part, full[kp])

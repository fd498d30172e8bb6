"""Where the time of the end-to-end API calls goes (tuning only): cProfile of a cold SHOT call and of a hand-off match."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from shot_fpfh_b200.descriptors import ShotMultiprocessor  # noqa: E402
from shot_fpfh_b200.matching import basic_matching  # noqa: E402

pts, normals, kp, radius = bench.make_shot_workload(0)
kept = []
with ShotMultiprocessor(min_neighborhood_size=10, verbose=False) as shot:
    for i in range(3):
        kept.append(shot.compute_descriptor_single_scale(pts.copy(), normals.copy(), kp.copy(), radius))
    c_pts, c_nrm, c_kp = pts.copy(), normals.copy(), kp.copy()
    torch.cuda.synchronize()
    prof = cProfile.Profile()
    t0 = time.perf_counter()
    prof.enable()
    kept.append(shot.compute_descriptor_single_scale(c_pts, c_nrm, c_kp, radius))
    prof.disable()
    print("cold SHOT call: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
    pstats.Stats(prof).sort_stats("cumulative").print_stats(18)
    if len(sys.argv) > 1:
        q = 200_000
        from shot_fpfh_b200 import synthetic

        scan, nrm = synthetic.bumpy_sphere(1_000_000, seed=0)
        ref, ref_nrm, _, _, _ = synthetic.rigid_pair(scan, nrm)
        d_scan = shot.compute_descriptor_single_scale(scan, nrm, scan[::5][:q], radius)
        d_ref = shot.compute_descriptor_single_scale(ref, ref_nrm, ref[::5][:q], radius)
        basic_matching(d_scan, d_ref)
        torch.cuda.synchronize()
        prof = cProfile.Profile()
        t0 = time.perf_counter()
        prof.enable()
        basic_matching(d_scan, d_ref)
        prof.disable()
        print("hand-off match: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
        pstats.Stats(prof).sort_stats("cumulative").print_stats(22)

"""
BASELINE.json configs[4] ("C5"): the UNMODIFIED reference's own orchestration — `shot_fpfh.pipeline.RegistrationPipeline`
from baseline/_ref, driven exactly as `scripts/register_point_clouds.py:80-127` drives it (select_keypoints ->
compute_descriptors -> find_descriptors_matches -> run_ransac -> run_icp) — on a synthetic rigid pair, once on top of
this package (`dropin.install()`, one B200) and once as it is (CPU), per-stage wall clock and the error of the recovered
transform against the known one.

    python scripts/run_reference_pipeline.py --points 10000000                 # B200 leg
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/run_reference_pipeline.py --points 10000000                    # 8 B200s: rank 0 runs the reference's
                                                                               # pipeline, ranks 1..7 lend their GPUs
                                                                               # (shot_fpfh_b200.distributed.serve)
    python scripts/run_reference_pipeline.py --points 300000 --cpu             # the reference alone, decimated (8d)

The pipeline object, its methods, their arguments and every line of pipeline.py are the reference's; only the callables
it imports are rebound.
"""

import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n: int, on_gpu: bool, fpfh: bool = False, verbose: bool = True) -> dict:
    from baseline import reference_loader
    from shot_fpfh_b200 import synthetic  # (the generator only: NumPy)

    reference_loader.load()
    if on_gpu:
        import torch

        import shot_fpfh_b200.dropin as dropin
        import shot_fpfh_b200.matching.ransac as ransac_module

        dropin.install()
        ransac_module.rng = np.random.default_rng(seed=72)
        sync = torch.cuda.synchronize
    else:
        sync = lambda: None  # noqa: E731
    pipeline = importlib.import_module("shot_fpfh.pipeline")
    times = {}

    def stage(name, fn):
        sync()
        t0 = time.perf_counter()
        out = fn()
        sync()
        times[name] = time.perf_counter() - t0
        if verbose:
            print(f"  {name:28s} {1e3 * times[name]:12.1f} ms", flush=True)
        return out

    scan, directions = synthetic.bumpy_sphere(n, seed=0)
    normals = synthetic.bumpy_sphere_true_normals(directions)
    ref, ref_normals, _, rot, trans = synthetic.rigid_pair(scan, normals)
    s = synthetic.mean_spacing(n)
    if verbose:
        print(f"{'B200 (dropin)' if on_gpu else 'reference alone (CPU)'}: pair of {n} points, radius {5 * s:.5f}", flush=True)
    pipe = pipeline.RegistrationPipeline(scan=scan, scan_normals=normals, ref=ref, ref_normals=ref_normals)
    stage("select_keypoints", lambda: pipe.select_keypoints("subsampling", neighborhood_size=3.75 * s))
    stage("compute_descriptors (SHOT)", lambda: pipe.compute_descriptors(
        radius=5.0 * s, descriptor_choice="shot_single_scale", subsample_support=False, min_neighborhood_size=10,
        n_procs=min(os.cpu_count() or 1, 64), disable_progress_bars=True, verbose=False))
    stage("find_descriptors_matches", lambda: pipe.find_descriptors_matches(
        "threshold", reject_threshold=0.8, threshold_multiplier=4.0))
    coarse, ratio = stage("run_ransac (10000 draws)", lambda: pipe.run_ransac(
        n_draws=10_000, max_inliers_distance=4 * s, disable_progress_bar=True))
    fine, rms, _ = stage("run_icp (point to plane)", lambda: pipe.run_icp(
        "point_to_plane", coarse, d_max=2 * s, voxel_size=3.75 * s, max_iter=20, rms_threshold=1e-9,
        disable_progress_bar=True))
    out = {
        "leg": "B200, dropin" if on_gpu else "reference alone, CPU", "n_points": n,
        "keypoints": [int(pipe.scan_keypoints.shape[0]), int(pipe.ref_keypoints.shape[0])],
        "matches": int(pipe.matches[0].shape[0]), "ransac_inlier_ratio": float(ratio),
        "coarse_rotation_error": float(np.abs(np.asarray(coarse.rotation) - rot).max()),
        "fine_rotation_error": float(np.abs(np.asarray(fine.rotation) - rot).max()),
        "fine_translation_error": float(np.abs(np.asarray(fine.translation) - trans).max()),
        "seconds": times, "total_seconds": float(sum(times.values())), "host_cores": os.cpu_count(),
    }
    if fpfh:
        pipe.scan_descriptors = pipe.ref_descriptors = None
        stage("compute_descriptors (FPFH)", lambda: pipe.compute_descriptors(
            radius=5.0 * s, descriptor_choice="fpfh", fpfh_n_bins=5, disable_progress_bars=True, verbose=False))
        out["seconds"] = times
    if on_gpu:
        import torch

        out["peak_device_GiB"] = torch.cuda.max_memory_allocated() / 2**30
    if verbose:
        print({k: v for k, v in out.items() if k != "seconds"}, flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--cpu", action="store_true", help="the reference alone (no dropin)")
    ap.add_argument("--fpfh", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not a.cpu:  # root + workers (the reference's script is a single process: it runs on rank 0)
        import torch
        import torch.distributed as dist

        from shot_fpfh_b200 import distributed as sfd

        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        warm = torch.ones(1 << 20, device="cuda")  # the communicator and its NVLink buffers exist before the clock starts
        dist.all_reduce(warm)
        dist.broadcast(warm, src=0)
        dist.broadcast_object_list([{"op": "warm"}], src=0)
        torch.cuda.synchronize()
        if dist.get_rank() != 0:
            served = sfd.serve()
            print(f"rank {dist.get_rank()}: served {served} requests", flush=True)
            dist.destroy_process_group()
            sys.exit(0)
        sfd.start_root_service()
        result = run(a.points, on_gpu=True, fpfh=a.fpfh)
        result["leg"] = f"{world} x B200: rank 0 runs the reference's pipeline (dropin), the other ranks serve the matching"
        sfd.stop_root_service()
        dist.destroy_process_group()
    else:
        result = run(a.points, on_gpu=not a.cpu, fpfh=a.fpfh)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(result, f, indent=1)

"""
Benchmark of the hot path on B200 (contract: see the task statement; summary in DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--skip-extra]

Headline workload (BASELINE.json configs[1], "C2"): SHOT single-scale on a seeded 1M-point synthetic surface scan,
~100k grid-selected query points, radius = 5 x mean spacing, min_neighborhood_size = 10.
One step = one pass of the whole entry point over one batch: grid build -> radius search (count, scan, fill)
-> local reference frames -> 352-bin descriptors.
  value : descriptors/s, inputs already resident in HBM, timed with CUDA events per step (L2 flushed between steps)
  e2e   : the same through the reference-shaped API `ShotMultiprocessor.compute_descriptor_single_scale` with HOST
          float64 arrays in pinned memory: H2D of cloud/normals/keypoints and D2H of the (Q, 352) float64 result
          are inside the timed region
  roofline    : dominant kernel, algorithmic bytes (SURVEY.md §8d formulas) / its CUDA-event duration vs measured HBM
  cpu_baseline: the oracle port (NumPy restatement of the reference, multiprocessing like the reference) on a bounded
                sample of the same workload on this host's cores
`--impl reference` times that CPU arm alone (rank 0 only under torchrun).
Multi-GPU (torchrun, one rank per GPU): queries shard by blocks with a replicated cloud and no data-path
collective; each rank processes a full-size block (weak scaling), value = all ranks' descriptors / max time.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 1_000_000
QUERY_VOXEL_IN_SPACINGS = 3.75
RADIUS_IN_SPACINGS = 5.0
MIN_NB = 10
OWN_KERNELS_PER_SHOT_STEP = 9  # bbox_init, bbox, key, place, rank_reorder, candidate_count, search_moments, lrf_eigen, shot_descriptor


_JSON_FD = None


def _quiet_stdout() -> None:
    """Everything any library prints to stdout during the run (NCCL's version banner under torchrun, progress bars)
    goes to stderr; the ONE JSON line is written to the original stdout by `_emit`."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=40)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--skip-extra", action="store_true", help="only the headline SHOT workload")
    p.add_argument("--multi-extra", action="store_true",
                   help="N > 1: also time the sharded FPFH and matching paths (the ones with an all-gather)")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the CPU baseline sample")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "tflops": float(d["bf16_tflops"]),
                "tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


def make_shot_workload(rank: int = 0):
    from shot_fpfh_b200 import synthetic

    pts, normals = synthetic.bumpy_sphere(N_POINTS, seed=0)
    s = synthetic.mean_spacing(N_POINTS)
    kp_idx = synthetic.voxel_first_point_queries(pts, QUERY_VOXEL_IN_SPACINGS * s)
    if rank:  # weak scaling: every rank gets its own full-size block of queries on the replicated cloud
        kp_idx = np.sort(np.random.default_rng(100 + rank).choice(N_POINTS, kp_idx.shape[0], replace=False))
    return pts, normals, np.ascontiguousarray(pts[kp_idx]), RADIUS_IN_SPACINGS * s


# ---------------------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------------------
def cpu_shot_sample(pts, normals, kp, radius, seconds: float, calib: dict | None = None):
    """
    Oracle port on a bounded sample of about `seconds` of CPU work. Returns (descriptors/s extrapolated to the full
    query set, info, calibration to reuse for the next sample).
    """
    from sklearn.neighbors import KDTree

    from oracle import shot_oracle

    cores = os.cpu_count() or 1
    n_procs = max(1, min(cores, 64))
    rng = np.random.default_rng(0)
    if calib is None:
        t0 = time.perf_counter()
        KDTree(pts)  # the reference builds the tree once per call (shot_parallelization.py:167)
        t_tree = time.perf_counter() - t0
        n_cal = min(kp.shape[0], 250 * n_procs)
        t0 = time.perf_counter()
        shot_oracle.shot_single_scale_pool(pts, normals, kp[rng.choice(kp.shape[0], n_cal, replace=False)], radius, True,
                                           MIN_NB, n_procs)
        t_cal = max(time.perf_counter() - t0 - t_tree, 1e-6)  # the pool driver builds the tree itself
        calib = {"t_tree": t_tree, "rate": n_cal / t_cal, "n_cal": n_cal}
    t_tree = calib["t_tree"]
    n_sample = int(min(kp.shape[0], max(calib["n_cal"], calib["rate"] * max(seconds - t_tree, 0.5))))
    t0 = time.perf_counter()
    shot_oracle.shot_single_scale_pool(pts, normals, kp[rng.choice(kp.shape[0], n_sample, replace=False)], radius, True,
                                       MIN_NB, n_procs)
    t_sample = max(time.perf_counter() - t0 - t_tree, 1e-6)
    full_time = t_tree + kp.shape[0] * t_sample / n_sample
    info = {
        "cores": n_procs,
        "kind": "port",
        "sample": (
            f"oracle port (NumPy restatement of shot_parallelization.py:135-183, multiprocessing.Pool of {n_procs}) on "
            f"{n_sample} of {kp.shape[0]} queries of the same 1M-point cloud: {t_sample:.2f} s + KDTree build "
            f"{t_tree:.2f} s; value = Q / (tree + Q * per-query time)"
        ),
    }
    return kp.shape[0] / full_time, info, calib


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pts, normals, kp, radius = make_shot_workload()
    # every step is a bounded sample; the whole run stays within ~2.5 minutes whatever K and W are
    per_step = min(8.0, args.cpu_seconds, max(1.5, 150.0 / max(1, args.warmup + args.steps)))
    vals, calib, info = [], None, None
    for i in range(args.warmup + args.steps):
        v, info, calib = cpu_shot_sample(pts, normals, kp, radius, per_step, calib)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference",
        "metric": "SHOT descriptors/sec",
        "value": value,
        "unit": "descriptors/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * kp.shape[0] / value,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"C2: SHOT single-scale, 1M-point synthetic surface scan, {kp.shape[0]} queries per GPU, radius 5x mean spacing",
                   "n_points": N_POINTS, "queries_per_gpu": int(kp.shape[0]), "min_neighborhood_size": MIN_NB},
        "cpu_baseline": {"value": value, "unit": "descriptors/s", **info},
        "e2e": {"value": value, "unit": "descriptors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={device_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=self.file, stderr=subprocess.DEVNULL,
            )
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        self.file.flush()
        self.file.seek(0)
        sm, mx, reasons = [], [], set()
        for row in self.file.read().strip().splitlines():
            f = [c.strip() for c in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.file.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


def timed_steps(step_fn, steps: int, warmup: int, flush, dist=None, after_step=None):
    """W untimed steps, then K steps each bracketed by CUDA events (L2 flushed, untimed, in between).
    `after_step` runs after the closing event of a step has been recorded (it may synchronise).
    Returns (ms per step = max over ranks of the mean, list of per-stage event dicts)."""
    import torch

    for _ in range(warmup):
        flush()
        step_fn(None)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    marks = []
    pairs = []
    for _ in range(steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage = {}
        a.record()
        step_fn(stage)
        b.record()
        if after_step is not None:
            after_step()
        pairs.append((a, b))
        marks.append(stage)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = float(np.mean([a.elapsed_time(b) for a, b in pairs]))
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    stages = {}
    for st in marks:
        for name, (e0, e1) in st.items():
            stages.setdefault(name, []).append(e0.elapsed_time(e1))
    return ms, {k: float(np.mean(v)) for k, v in stages.items()}


def mark(stage, name):
    """Context manager recording a CUDA-event pair on the current stream when `stage` is a dict."""
    import contextlib

    import torch

    @contextlib.contextmanager
    def cm():
        if stage is None:
            yield
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        stage[name] = (e0, e1)

    return cm()


def bench_shot(args, dist, rank, world, pk):
    import torch

    from shot_fpfh_b200 import ops
    from shot_fpfh_b200.descriptors import ShotMultiprocessor
    from shot_fpfh_b200.device import Grid, host_threads, upload

    pts, normals, kp, radius = make_shot_workload(rank)
    n, q = pts.shape[0], kp.shape[0]
    dev = torch.device("cuda")
    p_dev, n_dev, k_dev = upload(pts), upload(normals), upload(kp)
    grid = Grid()
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 4x the 126 MB L2
    out = torch.empty((q, 352), dtype=torch.float32, device=dev)
    counts = {}

    def flush():
        flush_buf.fill_(1)

    ops.profile_enable(True)
    kernel_ms = []

    def step(stage):
        with mark(stage, "grid_build"):
            grid.build(p_dev, n_dev, radius)
        # the fused single-scale driver: what ShotMultiprocessor.compute_descriptor_single_scale runs
        _, _, pairs = ops.shot_single_scale(grid, k_dev, radius, MIN_NB, True, out=out, want_pairs="pairs" not in counts)
        if pairs is not None:
            counts["pairs"] = pairs

    def read_kernel_events():  # CUDA events recorded by the driver around its three kernels; outside the bracket
        kernel_ms.append(ops.profile_read())

    sampler = ClockSampler(torch.cuda.current_device())
    ms, stages = timed_steps(step, args.steps, args.warmup, flush, dist, after_step=read_kernel_events)
    ops.profile_enable(False)
    k_ms = np.mean(np.array(kernel_ms), axis=0)
    stages.update({"search_moments": float(k_ms[0]), "lrf_eigen": float(k_ms[1]), "votes_descriptor": float(k_ms[2])})
    pairs = counts["pairs"]
    nonzero_rows = int((out.abs().sum(dim=1) > 0).sum().item())
    value = world * q / (ms * 1e-3)

    # algorithmic bytes per launch (float32 payloads, int32 indices)
    # B_search = 12N + 12Q + 4P + 4(Q+1), B_lrf = 16P + 48Q, B_shot = 28P + 48Q + 1408Q. The fused
    # driver finds the neighbours and accumulates the frame's moments in one kernel (B_search + the 48Q frame
    # scratch) and runs the frame's sign votes inside the descriptor kernel (B_shot + the 16P gather of B_lrf).
    alg = {
        "grid_build": 52 * n,
        "search_moments": 12 * n + 12 * q + 4 * pairs + 4 * (q + 1) + 48 * q,
        "lrf_eigen": 2 * 48 * q,
        "votes_descriptor": 28 * pairs + 48 * q + 4 * 352 * q + 16 * pairs,
    }
    dominant = max(stages, key=stages.get)
    achieved = alg[dominant] / (stages[dominant] * 1e-3) / 1e9
    traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of this workload
    traffic_file = os.path.join(ROOT, "profiles", "traffic_c2.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            traffic = json.load(f).get(dominant)
    roofline = {
        "kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
        "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["source"],
        "algorithmic_bytes": alg[dominant], "kernel_ms": stages[dominant],
        "per_stage": {k: {"ms": stages[k], "algorithmic_GBps": alg[k] / (stages[k] * 1e-3) / 1e9} for k in stages},
    }

    # ---- end to end through the reference-shaped API, host buffers in pinned memory ----
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()

    h_pts, h_nrm, h_kp = pinned(pts), pinned(normals), pinned(kp)
    e2e_times = []
    n_threads = host_threads(ShotMultiprocessor.n_procs)  # the reference's default worker count (8)
    with ShotMultiprocessor(min_neighborhood_size=MIN_NB, verbose=False) as shot:
        for i in range(args.warmup + args.steps):
            flush()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            d = shot.compute_descriptor_single_scale(h_pts, h_nrm, h_kp, radius)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                e2e_times.append(dt)
    clocks = sampler.stop()  # sampled over both timed regions (device-resident steps and end-to-end steps)
    assert d.shape == (q, 352) and d.dtype == np.float64
    e2e_ms = float(np.mean(e2e_times)) * 1e3
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {
        "value": world * q / (e2e_ms * 1e-3), "unit": "descriptors/s", "ms_per_step": e2e_ms,
        "h2d_bytes_per_step": int(h_pts.nbytes + h_nrm.nbytes + h_kp.nbytes), "d2h_bytes_per_step": int(shot.last_d2h_bytes),
        "host_threads": n_threads,
        "transport": "in blocks of queries: the rows (~86% zeros) are compacted on the device (offsets + uint16 column + float32 value per non-zero), copied, and expanded into the float64 result by the host threads while the next block is computed",
        "api": "ShotMultiprocessor.compute_descriptor_single_scale(point_cloud, normals, keypoints, radius) -> float64 (Q,352)",
    }
    config = {
        "workload": f"C2: SHOT single-scale, 1M-point synthetic surface scan, {q} queries per GPU, radius 5x mean spacing",
        "n_points": n, "queries_per_gpu": q, "neighbour_pairs": pairs, "mean_neighbours": pairs / q,
        "min_neighborhood_size": MIN_NB, "nonzero_rows": nonzero_rows, "l2": "flushed between steps (512 MB write)",
        "sharding": "queries by block, replicated cloud, no data-path collective" if world > 1 else "single GPU",
        "output": "float32 (Q,352) resident in HBM for `value`; float64 on the host for `e2e`",
    }
    grid.close()
    return {"ms": ms, "value": value, "roofline": roofline, "e2e": e2e, "config": config, "clocks": clocks,
            "host": (pts, normals, kp, radius), "launches": OWN_KERNELS_PER_SHOT_STEP * args.steps}


def _pinned(a):
    import torch

    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t.numpy()


def bench_fpfh(args, pk):
    """C3: FPFH 33-d on the full 1M-point cloud, every point a query."""
    import torch

    from shot_fpfh_b200 import ops, synthetic
    from shot_fpfh_b200.device import Grid, upload

    pts, normals = synthetic.bumpy_sphere(N_POINTS, seed=0)
    radius = RADIUS_IN_SPACINGS * synthetic.mean_spacing(N_POINTS)
    p_dev, n_dev = upload(pts), upload(normals)
    kp = torch.arange(N_POINTS, dtype=torch.int64, device="cuda")
    grid = Grid()
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    info = {}

    out = torch.empty((N_POINTS, 33), dtype=torch.float32, device="cuda")
    kernel_ms = []
    ops.profile_enable(True)

    def step(stage):
        with mark(stage, "grid_build"):
            grid.build(p_dev, n_dev, radius)
        # the fused driver: what compute_fpfh_descriptor runs (search + weights, SPFH, FPFH; padded neighbour list)
        _, pairs = ops.fpfh_cloud(grid, radius, 11, True, kp, out=out, want_pairs="pairs" not in info)
        if pairs is not None:
            info["pairs"] = pairs

    steps = max(3, args.steps // 2)
    ms, stages = timed_steps(step, steps, args.warmup, lambda: flush_buf.fill_(1),
                             after_step=lambda: kernel_ms.append(ops.profile_read()))
    ops.profile_enable(False)
    k_ms = np.mean(np.array(kernel_ms), axis=0)
    stages.update({"radius_search": float(k_ms[0]), "spfh": float(k_ms[1]), "fpfh": float(k_ms[2])})
    p, n, d = info["pairs"], N_POINTS, 33
    # SURVEY.md §8d; the search writes index + float32 weight per pair (8P) and the padded offsets / counts
    alg = {"grid_build": 52 * n, "radius_search": 12 * n + 12 * n + 8 * p + 12 * (n + 1), "spfh": 28 * p + 24 * n + 4 * d * n,
           "fpfh": (4 * d + 8) * p + 4 * d * n + 4 * n}
    dominant = max(stages, key=stages.get)
    traffic = None
    traffic_file = os.path.join(ROOT, "profiles", "traffic_c3.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            traffic = json.load(f).get(dominant)
    grid.close()
    # end to end through the reference-shaped call: host float64 arrays in (pinned), (N, 33) float64 host array out
    from shot_fpfh_b200.descriptors import compute_fpfh_descriptor

    h_pts, h_nrm = _pinned(pts), _pinned(normals)
    h_kp = np.arange(N_POINTS, dtype=np.int64)
    e2e_times = []
    for i in range(7):  # three untimed calls: the host result buffers settle at the third (device.result_buffer)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = compute_fpfh_descriptor(h_kp, h_pts, h_nrm, radius, n_bins=11, decorrelated=True, verbose=False)
        if i >= 3:
            e2e_times.append(time.perf_counter() - t0)
    e2e_ms = float(np.mean(e2e_times)) * 1e3
    return {
        "workload": "C3: FPFH 33-d (n_bins=11, decorrelated), 1M-point cloud, every point a query, 1 GPU",
        "value": n / (ms * 1e-3), "unit": "descriptors/s", "ms_per_step": ms, "steps": steps, "neighbour_pairs": p,
        "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "descriptors/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h_pts.nbytes + h_nrm.nbytes + h_kp.nbytes), "d2h_bytes_per_step": int(rows.nbytes // 2),
                "transport": "float32 rows cross PCIe by blocks of keypoints while the next block is computed; host threads widen them exactly into the float64 result",
                "api": "compute_fpfh_descriptor(keypoints_indices, cloud_points, normals, radius, n_bins=11, decorrelated=True) -> float64 (N,33)"},
        "roofline": {"kernel": dominant, "bound": "hbm", "achieved": alg[dominant] / (stages[dominant] * 1e-3) / 1e9,
                     "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": alg[dominant] / (stages[dominant] * 1e-3) / 1e9 / pk["hbm_gbs"],
                     "traffic": traffic, "algorithmic_bytes": alg[dominant],
                     "note": "algorithmic bytes count one SPFH-row gather per neighbour pair (SURVEY.md 8d); the 144 MB table "
                             "is served by L1/L2, so the figure on algorithmic bytes can exceed the HBM peak: see `traffic` "
                             "(DRAM bytes per launch, ncu) - the FPFH stage is bound by the L1 data pipe, not by HBM",
                     "per_stage": {k: {"ms": stages[k], "algorithmic_GBps": alg[k] / (stages[k] * 1e-3) / 1e9} for k in stages}},
    }


def bench_registration(args, pk):
    """
    The stages after the matcher, through the reference-shaped API with host arrays in and
    out: `ransac_on_matches` (10 000 draws over 20 000 matches, a third of them wrong) and `icp_point_to_plane`
    (1M-point pair, ~100k voxel-subsampled scan points, 10 iterations). CPU figures: the oracle on a bounded part
    of the same work (RANSAC: 100 draws, ICP: tree build + 2 iterations), scaled and labelled.
    """
    import shot_fpfh_b200.matching.ransac as ransac
    from oracle import registration_oracle as ro
    from shot_fpfh_b200 import synthetic
    from shot_fpfh_b200.core import RigidTransform
    from shot_fpfh_b200.icp import icp_point_to_plane

    scan, normals = synthetic.bumpy_sphere(N_POINTS, seed=0)
    ref, ref_normals, perm, rot, trans = synthetic.rigid_pair(scan, normals)
    s = synthetic.mean_spacing(N_POINTS)
    true_ref = np.empty(N_POINTS, dtype=np.int64)
    true_ref[perm] = np.arange(N_POINTS)
    rng = np.random.default_rng(5)
    m = 20_000
    scan_idx = rng.choice(N_POINTS, m, replace=False)
    ref_idx = true_ref[scan_idx].copy()
    wrong = rng.random(m) < 0.33
    ref_idx[wrong] = rng.integers(0, N_POINTS, int(wrong.sum()))
    out = {"workload": f"RANSAC 10000 draws x {m} matches; point-to-plane ICP, 1M-point pair, 10 iterations"}
    times = []
    for _ in range(3):
        ransac.rng = np.random.default_rng(seed=72)
        t0 = time.perf_counter()
        ratio, best = ransac.ransac_on_matches(scan_idx, ref_idx, scan, ref, n_draws=10_000, distance_threshold=2 * s)
        times.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    ransac.replay_choices(ransac.rng, m, 4, 10_000)  # NumPy's draw stream, replayed in one vectorised pass
    t_draws = time.perf_counter() - t0
    t0 = time.perf_counter()
    ro.ransac_on_matches(scan_idx, ref_idx, scan, ref, np.random.default_rng(seed=72), n_draws=100, distance_threshold=2 * s)
    cpu_ransac = (time.perf_counter() - t0) * 100
    out["ransac"] = {"ms": min(times) * 1e3, "of_which_host_draw_replay_ms": t_draws * 1e3, "inlier_ratio": ratio,
                     "cpu_oracle_ms_extrapolated_from_100_draws": cpu_ransac * 1e3}
    init = RigidTransform(best.rotation, best.translation)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        t_icp, rms, _ = icp_point_to_plane(scan, ref, ref_normals, init, d_max=2 * s, voxel_size=QUERY_VOXEL_IN_SPACINGS * s,
                                           max_iter=10, rms_threshold=0.0)
        times.append(time.perf_counter() - t0)
    from shot_fpfh_b200.subsampling import grid_subsampling

    sub = grid_subsampling(scan, QUERY_VOXEL_IN_SPACINGS * s)
    t0 = time.perf_counter()
    ro.icp_point_to_plane(scan, ref, ref_normals, (best.rotation, best.translation), 2 * s, sub, max_iter=2, rms_threshold=0.0)
    cpu_icp2 = time.perf_counter() - t0
    out["icp_point_to_plane"] = {"ms": min(times) * 1e3, "iterations": 10, "subsampled_points": int(sub.shape[0]),
                                 "final_mean_residual": float(rms),
                                 "rotation_error_max_abs": float(np.abs(t_icp.rotation - rot).max()),
                                 "cpu_oracle_ms_tree_build_plus_2_iterations": cpu_icp2 * 1e3}
    return out


def bench_match(args, pk, q: int = 200_000):
    """C4: 200k x 200k 352-d exact NN (+ second NN for the ratio test): shortlist GEMM on tensor cores + fp64 re-rank."""
    import torch

    from shot_fpfh_b200 import ops, synthetic

    a = torch.from_numpy(synthetic.sparse_unit_rows(q, 352, seed=2)).cuda().double()
    b = torch.from_numpy(synthetic.sparse_unit_rows(q, 352, seed=3)).cuda().double()
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def step(stage):
        with mark(stage, "nonempty_pack"):
            ra, rb = ops.nonempty_rows(a), ops.nonempty_rows(b)
            ap, _ = ops.match_pack(a, ra, 1.0)
            bp, bn = ops.match_pack(b, rb, 1.0)
        with mark(stage, "shortlist_gemm"):
            _, cand = ops.match_topk(ap, bp, bn, 8, 0, True)
        with mark(stage, "rerank_f64"):
            ops.match_rerank(a, ra, b, rb, cand)

    steps = max(3, args.steps // 3)
    ms, stages = timed_steps(step, steps, min(args.warmup, 3), lambda: flush_buf.fill_(1))
    flops = 2.0 * q * q * 352
    tf = flops / (stages["shortlist_gemm"] * 1e-3) / 1e12
    # end to end through the reference-shaped call: two (q, 352) float64 host arrays in (pinned), index pairs out
    from shot_fpfh_b200.matching import basic_matching

    h_a, h_b = _pinned(a.cpu().numpy()), _pinned(b.cpu().numpy())
    e2e_times = []
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pairs = basic_matching(h_a, h_b)
        if i >= 1:
            e2e_times.append(time.perf_counter() - t0)
    e2e_ms = float(np.mean(e2e_times)) * 1e3
    return {
        "workload": f"C4: {q} x {q} x 352 exact nearest + second-nearest neighbour, synthetic sparse unit rows, 1 GPU",
        "value": q / (ms * 1e-3), "unit": "match queries/s", "ms_per_step": ms, "steps": steps,
        "e2e": {"value": q / (e2e_ms * 1e-3), "unit": "match queries/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h_a.nbytes + h_b.nbytes), "d2h_bytes_per_step": int(pairs[0].nbytes + pairs[1].nbytes),
                "api": "basic_matching(scan_descriptors, ref_descriptors)"},
        "roofline": {"kernel": "topk_tc_kernel", "bound": "tensor", "achieved": tf, "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": tf / pk["tflops"], "flops": flops,
                     "per_stage_ms": stages},
    }


def bench_distributed(args, dist, rank, world):
    """
    N > 1 only: the two sharded paths that DO exchange data, strong scaling of one job over the ranks,
    through `shot_fpfh_b200.distributed` with host arrays in: FPFH C3 (SPFH by blocks of the cell-sorted cloud, ONE
    all-gather of the SPFH rows over NVLink, FPFH by keypoint blocks) and matching (target set sharded, every rank
    emits its exact nearest / second nearest, ONE all-gather of a (Q, 3) float64 tensor, merge). Wall clock between
    barriers, max over ranks, best of 3.
    """
    import torch

    from shot_fpfh_b200 import distributed, synthetic

    out = {}
    pts, normals = synthetic.bumpy_sphere(N_POINTS, seed=0)
    radius = RADIUS_IN_SPACINGS * synthetic.mean_spacing(N_POINTS)
    h_pts, h_nrm = _pinned(pts), _pinned(normals)
    kp = np.arange(N_POINTS, dtype=np.int64)

    def timed(fn, reps=3):
        best = float("inf")
        for _ in range(reps + 1):  # first call warms the pools
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best * 1e3

    ms = timed(lambda: distributed.fpfh(kp, h_pts, h_nrm, radius, 11, decorrelated=True, gather=False))
    out["fpfh_c3_sharded"] = {
        "workload": f"C3 over {world} GPUs: every rank uploads the cloud and builds the grid, SPFH of its block of the "
                    "cell-sorted cloud, one all-gather of SPFH rows (132 MB in all), FPFH of its block of keypoints",
        "ms": ms, "value": N_POINTS / (ms * 1e-3), "unit": "descriptors/s", "scaling": "strong",
    }
    qm = 200_000
    a = _pinned(synthetic.sparse_unit_rows(qm, 352, seed=2).astype(np.float64))
    b = _pinned(synthetic.sparse_unit_rows(qm, 352, seed=3).astype(np.float64))
    ms = timed(lambda: distributed.nearest_neighbors(a, b), reps=2)
    out["match_c4_sharded"] = {
        "workload": f"C4 over {world} GPUs: {qm} x {qm} x 352, target set sharded, every rank uploads both sets "
                    "(1.13 GB of float64 rows over its own PCIe link), one all-gather of (Q, 3) float64, merge",
        "ms": ms, "value": qm / (ms * 1e-3), "unit": "match queries/s", "scaling": "strong",
    }
    return out


def main():
    args = parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk = peaks()
    res = bench_shot(args, dist, rank, world, pk)
    extra = {}
    cpu = None
    if rank == 0 and world == 1:
        pts, normals, kp, radius = res["host"]
        v, info, _ = cpu_shot_sample(pts, normals, kp, radius, args.cpu_seconds)
        cpu = {"value": v, "unit": "descriptors/s", **info}
        if not args.skip_extra:
            for name, fn in (("fpfh_c3", bench_fpfh), ("match_c4", bench_match), ("registration", bench_registration)):
                try:
                    extra[name] = fn(args, pk)
                except Exception as exc:  # noqa: BLE001  (the headline line must still be printed)
                    extra[name] = {"error": f"{type(exc).__name__}: {exc}"}
    if world > 1 and args.multi_extra:
        try:
            extra["multi_gpu"] = bench_distributed(args, dist, rank, world)
        except Exception as exc:  # noqa: BLE001
            extra["multi_gpu"] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        line = {
            "metric": "SHOT descriptors/sec",
            "value": res["value"],
            "unit": "descriptors/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": res["ms"],
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": res["config"],
            "roofline": res["roofline"],
            "cpu_baseline": cpu,
            "e2e": res["e2e"],
            "gpu_launches": res["launches"],
            "clocks": res["clocks"],
            "extra": extra,
        }
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""
Benchmark of the hot path on B200 (contract: see the task statement; summary in DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--skip-extra]

Headline workload (BASELINE.json configs[1], "C2"): SHOT single-scale on a seeded 1M-point synthetic surface scan,
~100k grid-selected query points, radius = 5 x mean spacing, min_neighborhood_size = 10.
One step = one pass of the whole entry point over one batch: grid build -> candidate count + scan -> search + frame
moments -> frame eigensolve -> 352-bin descriptors (float32-filtered kernel, float64 kernel on what it hands over).
  value : descriptors/s, inputs already resident in HBM, timed with CUDA events per step (L2 flushed between steps);
          the handle runs in its no-host-synchronisation mode (assumptions checked on the device, polled at the end)
  e2e   : the same through the reference-shaped API `ShotMultiprocessor.compute_descriptor_single_scale` with HOST
          float64 arrays in pinned memory: H2D of cloud/normals/keypoints and D2H of the (Q, 352) float64 result
          are inside the timed region; `e2e.cold` is the same call with fresh PAGEABLE arrays and every result kept
  roofline    : dominant kernel, algorithmic bytes (SURVEY.md 8d formulas) / its CUDA-event duration vs measured HBM
  cpu_baseline: the UNMODIFIED reference (baseline/_ref, `ShotMultiprocessor` with a multiprocessing.Pool over all the
                host's cores) on a bounded sample of the same workload; the oracle port when baseline/_ref is absent
`--impl reference` times that CPU arm alone (rank 0 only under torchrun).
`extra`: C3 (FPFH 33-d, 1M points) and C4 (200k x 200k matching, synthetic and real SHOT rows), each with roofline, e2e
and its own cpu_baseline, and the registration stages. N > 1 (torchrun, one rank per GPU): the headline is weak scaling
(queries shard by blocks, replicated cloud, no data-path collective; value = all ranks' descriptors / max time) and
`extra.multi_gpu` times the two paths that DO exchange data, strong scaling of one job: C3 (one all-gather of SPFH
rows) and C4 (target set sharded, one all-gather of per-rank results), with the collective's share.
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
# own kernels per steady-state step: key, place, rank_reorder (grid); candidate_count, search_moments, lrf_eigen,
# shot_fast, shot_descriptor (the float64 kernel on the handed-over queries) — the cell-table prefix sum is CUB's
OWN_KERNELS_PER_SHOT_STEP = 8


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
    p.add_argument("--multi-extra", action="store_true", help="(kept for old command lines: the sharded legs now always run)")
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
def shot_config(q: int) -> dict:
    """The workload description both arms print (identical dicts: the driver compares them)."""
    return {
        "workload": f"C2: SHOT single-scale, 1M-point synthetic surface scan, {q} queries per GPU, radius 5x mean spacing",
        "n_points": N_POINTS, "queries_per_gpu": int(q), "min_neighborhood_size": MIN_NB,
        "l2": "flushed between steps (512 MB write)",
    }


def _reference():
    """The unmodified reference from baseline/_ref, or None (then the oracle port stands in)."""
    try:
        from baseline import reference_loader
    except ImportError:
        return None
    if not reference_loader.available():
        return None
    return reference_loader.load()


def _cpu_shot(pts, normals, kp, radius, n_procs):
    """One CPU call: the reference's ShotMultiprocessor when it is here, else the oracle port. -> kind"""
    ref = _reference()
    if ref is not None:
        import importlib

        descriptors = importlib.import_module("shot_fpfh.descriptors")
        with descriptors.ShotMultiprocessor(normalize=True, min_neighborhood_size=MIN_NB, n_procs=n_procs,
                                            disable_progress_bar=True, verbose=False) as shot:
            shot.compute_descriptor_single_scale(pts, normals, kp, radius)
        return "reference"
    from oracle import shot_oracle

    shot_oracle.shot_single_scale_pool(pts, normals, kp, radius, True, MIN_NB, n_procs)
    return "port"


def cpu_shot_sample(pts, normals, kp, radius, seconds: float, calib: dict | None = None):
    """
    The CPU path on a bounded sample of about `seconds` of work. Returns (descriptors/s extrapolated to the full query
    set, info, calibration to reuse for the next sample).
    """
    from sklearn.neighbors import KDTree

    cores = os.cpu_count() or 1
    n_procs = max(1, min(cores, 64))
    rng = np.random.default_rng(0)
    if calib is None:
        t0 = time.perf_counter()
        KDTree(pts)  # the reference builds the tree once per call (shot_parallelization.py:167)
        t_tree = time.perf_counter() - t0
        n_cal = min(kp.shape[0], 250 * n_procs)
        t0 = time.perf_counter()
        _cpu_shot(pts, normals, kp[rng.choice(kp.shape[0], n_cal, replace=False)], radius, n_procs)
        t_cal = max(time.perf_counter() - t0 - t_tree, 1e-6)  # the call builds the tree itself
        calib = {"t_tree": t_tree, "rate": n_cal / t_cal, "n_cal": n_cal}
    t_tree = calib["t_tree"]
    n_sample = int(min(kp.shape[0], max(calib["n_cal"], calib["rate"] * max(seconds - t_tree, 0.5))))
    t0 = time.perf_counter()
    kind = _cpu_shot(pts, normals, kp[rng.choice(kp.shape[0], n_sample, replace=False)], radius, n_procs)
    t_sample = max(time.perf_counter() - t0 - t_tree, 1e-6)
    full_time = t_tree + kp.shape[0] * t_sample / n_sample
    what = ("the unmodified reference (baseline/_ref): ShotMultiprocessor.compute_descriptor_single_scale, "
            f"multiprocessing.Pool of {n_procs}" if kind == "reference" else
            f"oracle port (NumPy restatement of shot_parallelization.py:135-183, multiprocessing.Pool of {n_procs})")
    info = {
        "cores": n_procs,
        "kind": kind,
        "sample": (
            f"{what} on {n_sample} of {kp.shape[0]} queries of the same 1M-point cloud: {t_sample:.2f} s + KDTree build "
            f"{t_tree:.2f} s; value = Q / (tree + Q * per-query time)"
        ),
    }
    return kp.shape[0] / full_time, info, calib


def cpu_fpfh_sample(seconds: float) -> dict:
    """
    FPFH on the CPU (single-threaded by construction, fpfh.py:38-116): a smaller cloud of the same generator with its own
    radius = 5 x spacing, i.e. the same ~72 neighbours per point as C3, every point a query; value = points / s.
    The 33-d layout the GPU leg computes raises in the unmodified reference (fpfh.py:59-79), so the 33-d figure is the
    oracle port's (reference + the one-token fix); the reference's own 125-d layout is timed beside it when it is here.
    """
    from oracle import fpfh_oracle
    from shot_fpfh_b200 import synthetic

    n = 6000
    pts, normals = synthetic.bumpy_sphere(n, seed=0)
    radius = RADIUS_IN_SPACINGS * synthetic.mean_spacing(n)
    t0 = time.perf_counter()
    fpfh_oracle.fpfh(np.arange(n), pts, normals, radius, 11, True)
    rate = n / (time.perf_counter() - t0)
    n = int(min(60_000, max(6000, rate * seconds)))
    pts, normals = synthetic.bumpy_sphere(n, seed=0)
    radius = RADIUS_IN_SPACINGS * synthetic.mean_spacing(n)
    t0 = time.perf_counter()
    fpfh_oracle.fpfh(np.arange(n), pts, normals, radius, 11, True)
    t = time.perf_counter() - t0
    out = {"value": n / t, "unit": "descriptors/s", "cores": 1, "kind": "port",
           "sample": f"oracle port of fpfh.py:16-117 (33-d, decorrelated, patched) on a {n}-point cloud of the same generator, "
                     f"every point a query, same neighbours per point as C3: {t:.2f} s"}
    if _reference() is not None:
        import importlib

        compute = importlib.import_module("shot_fpfh.descriptors").compute_fpfh_descriptor
        m = max(3000, n // 3)
        p2, n2 = synthetic.bumpy_sphere(m, seed=0)
        t0 = time.perf_counter()
        compute(np.arange(m), p2, n2, radius=RADIUS_IN_SPACINGS * synthetic.mean_spacing(m), n_bins=5,
                disable_progress_bars=True, verbose=False)
        out["reference_125d"] = {"value": m / (time.perf_counter() - t0), "unit": "descriptors/s", "kind": "reference",
                                 "sample": f"unmodified reference, n_bins=5 (125-d), {m}-point cloud, every point a query"}
    return out


def cpu_match_sample(a: np.ndarray, b: np.ndarray, q_full: int, seconds: float) -> dict:
    """
    `basic_matching` of the reference (cdist + argmin, matching.py:149-169) on m x m rows of the same sets, m sized for
    about `seconds`; the 200k x 200k figure is EXTRAPOLATED with the O(Q T) cost (the reference cannot allocate the
    320 GB matrix).
    """
    ref = _reference()
    if ref is not None:
        import importlib

        match, kind = importlib.import_module("shot_fpfh.matching").basic_matching, "reference"
    else:
        from oracle import matching_oracle

        match, kind = matching_oracle.basic_matching, "port"
    m = 1500
    t0 = time.perf_counter()
    match(a[:m], b[:m])
    pair_rate = m * m / (time.perf_counter() - t0)
    m = int(min(a.shape[0], b.shape[0], 20_000, max(1500, np.sqrt(pair_rate * seconds))))
    t0 = time.perf_counter()
    match(a[:m], b[:m])
    t = time.perf_counter() - t0
    t_full = t * (q_full / m) ** 2
    return {"value": q_full / t_full, "unit": "match queries/s", "cores": 1, "kind": kind, "extrapolated": True,
            "sample": f"basic_matching on {m} x {m} x {a.shape[1]} rows of the same sets: {t:.2f} s; scaled by (Q/m)^2 to "
                      f"{q_full} x {q_full} ({t_full:.0f} s; the reference cannot allocate that distance matrix)"}


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
        "config": shot_config(kp.shape[0]),
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
    # the handle in its no-host-synchronisation mode: after the first step the rebuilt grid assumes the same box and
    # the fused driver the same list size; both are checked on the device and polled after the timed loop
    grid = Grid().set_speculative(builds=True, shot_lists=True)
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 4x the 126 MB L2
    out = torch.empty((q, 352), dtype=torch.float32, device=dev)

    def flush():
        flush_buf.fill_(1)

    grid.build(p_dev, n_dev, radius)
    _, _, pairs = ops.shot_single_scale(grid, k_dev, radius, MIN_NB, True, out=out, want_pairs=True)
    handed_over = ops.shot_last_deferred()
    kernel_ms = []

    def step(stage):
        grid.build(p_dev, n_dev, radius)
        # the fused single-scale driver: what ShotMultiprocessor.compute_descriptor_single_scale runs
        ops.shot_single_scale(grid, k_dev, radius, MIN_NB, True, out=out)

    def step_with_stage_events(stage):
        with mark(stage, "grid_build"):
            grid.build(p_dev, n_dev, radius)
        ops.shot_single_scale(grid, k_dev, radius, MIN_NB, True, out=out)

    def read_kernel_events():  # CUDA events recorded by the driver around its three stages; outside the bracket
        kernel_ms.append(ops.profile_read())

    sampler = ClockSampler(torch.cuda.current_device())
    # (1) the timed region of `value`: K steps, one CUDA-event pair around each, nothing recorded inside a step
    ms, _ = timed_steps(step, args.steps, args.warmup, flush, dist)
    # (2) the same K steps again with an event between the stages (six records per step: they cost about 14 us of a
    #     0.5 ms step, which is why (1) does not carry them); the roofline's kernel time comes from this pass
    ops.profile_enable(True)
    ms_staged, stages = timed_steps(step_with_stage_events, args.steps, 1, flush, dist, after_step=read_kernel_events)
    ops.profile_enable(False)
    assert grid.poll() == 0, "a step without host synchronisation did nothing (its device-side check failed)"
    k_ms = np.mean(np.array(kernel_ms), axis=0)
    stages.update({"search_moments": float(k_ms[0]), "lrf_eigen": float(k_ms[1]), "descriptor": float(k_ms[2])})
    nonzero_rows = int((out.abs().sum(dim=1) > 0).sum().item())
    value = world * q / (ms * 1e-3)

    # algorithmic bytes per launch, SURVEY.md 8d (float32 payloads, int32 indices):
    #   B_grid = 52 N;  B_search = 12 N + 12 Q + 4 P + 4 (Q + 1), + 48 Q of frame moments written by the same kernel;
    #   B_eigen = 2 * 48 Q;  B_shot = 28 P + 48 Q + 4 * 352 Q  (the frame's sign votes run inside the descriptor kernel
    #   on the same neighbours: no bytes of their own)
    alg = {
        "grid_build": 52 * n,
        "search_moments": 12 * n + 12 * q + 4 * pairs + 4 * (q + 1) + 48 * q,
        "lrf_eigen": 2 * 48 * q,
        "descriptor": 28 * pairs + 48 * q + 4 * 352 * q,
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
        "whole_step": {"algorithmic_bytes": int(sum(alg.values())), "GBps": sum(alg.values()) / (ms * 1e-3) / 1e9,
                       "frac": sum(alg.values()) / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
        "per_stage": {k: {"ms": stages[k], "algorithmic_GBps": alg[k] / (stages[k] * 1e-3) / 1e9} for k in stages},
        "stage_times": f"CUDA events between the stages, over a second pass of {args.steps} steps ({ms_staged:.4f} ms per step "
                       "with the six extra event records; the steps timed for `value` record nothing inside a step)",
    }

    # ---- end to end through the reference-shaped API ----
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()

    h_pts, h_nrm, h_kp = pinned(pts), pinned(normals), pinned(kp)
    e2e_times, cold_times, kept = [], [], []
    n_threads = host_threads(ShotMultiprocessor.n_procs)  # the reference's default worker count (8)
    with ShotMultiprocessor(min_neighborhood_size=MIN_NB, verbose=False) as shot:
        for i in range(args.warmup + args.steps):  # (a) host buffers in pinned memory, the result dropped per call
            flush()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            d = shot.compute_descriptor_single_scale(h_pts, h_nrm, h_kp, radius)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                e2e_times.append(dt)
        d2h = int(shot.last_d2h_bytes)
        assert d.shape == (q, 352) and d.dtype == np.float64
        del d
        # (b) what a pipeline does (pipeline.py:154-174): fresh pageable NumPy arrays in, every result kept
        for i in range(2 + min(args.steps, 6)):
            c_pts, c_nrm, c_kp = pts.copy(), normals.copy(), kp.copy()
            flush()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            kept.append(shot.compute_descriptor_single_scale(c_pts, c_nrm, c_kp, radius))
            dt = time.perf_counter() - t0
            if i >= 2:
                cold_times.append(dt)
    clocks = sampler.stop()  # sampled over both timed regions (device-resident steps and end-to-end steps)
    del kept

    def over_ranks(values):
        v = float(np.mean(values)) * 1e3
        if dist is not None:
            t = torch.tensor([v], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            v = float(t.item())
        return v

    e2e_ms, cold_ms = over_ranks(e2e_times), over_ranks(cold_times)
    e2e = {
        "value": world * q / (e2e_ms * 1e-3), "unit": "descriptors/s", "ms_per_step": e2e_ms,
        "h2d_bytes_per_step": int(h_pts.nbytes + h_nrm.nbytes + h_kp.nbytes), "d2h_bytes_per_step": d2h,
        "host_threads": n_threads,
        "cold": {"value": world * q / (cold_ms * 1e-3), "ms_per_step": cold_ms,
                 "what": "the same call with fresh PAGEABLE NumPy arrays in and every result kept by the caller "
                         "(pipeline.py:154-174): the copies are staged by the driver, the 288 MB result is a new buffer"},
        "transport": "in blocks of queries: the rows (~86% zeros) are compacted on the device (offsets + uint16 column + float32 value per non-zero), copied, and expanded into the float64 result by the host threads while the next block is computed",
        "api": "ShotMultiprocessor.compute_descriptor_single_scale(point_cloud, normals, keypoints, radius) -> float64 (Q,352)",
    }
    config = shot_config(q)
    stats = {
        "neighbour_pairs": pairs, "mean_neighbours": pairs / q, "nonzero_rows": nonzero_rows,
        "queries_handed_to_the_float64_kernel": handed_over,
        "sharding": "queries by block, replicated cloud, no data-path collective" if world > 1 else "single GPU",
        "output": "float32 (Q,352) resident in HBM for `value`; float64 on the host for `e2e`",
        "host_synchronisations_per_step": 0,
    }
    grid.close()
    return {"ms": ms, "value": value, "roofline": roofline, "e2e": e2e, "config": config, "stats": stats, "clocks": clocks,
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
    # five untimed calls: `rows = f()` keeps the previous result alive during a call, so two host buffers alternate, and
    # each is replaced by a page-locked block the first time it is recycled (device.result_buffer): settled at the fifth
    for i in range(5 + max(4, min(args.steps, 8))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = compute_fpfh_descriptor(h_kp, h_pts, h_nrm, radius, n_bins=11, decorrelated=True, verbose=False)
        if i >= 5:
            e2e_times.append(time.perf_counter() - t0)
    e2e_ms = float(np.mean(e2e_times)) * 1e3
    try:
        cpu = cpu_fpfh_sample(min(args.cpu_seconds, 12.0))
    except Exception as exc:  # noqa: BLE001
        cpu = {"error": f"{type(exc).__name__}: {exc}"}
    whole = sum(alg.values())
    return {
        "cpu_baseline": cpu,
        "whole_step": {"algorithmic_bytes": int(whole), "GBps": whole / (ms * 1e-3) / 1e9, "frac_of_hbm": whole / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
        "workload": "C3: FPFH 33-d (n_bins=11, decorrelated), 1M-point cloud, every point a query, 1 GPU",
        "value": n / (ms * 1e-3), "unit": "descriptors/s", "ms_per_step": ms, "steps": steps, "neighbour_pairs": p,
        "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "descriptors/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h_pts.nbytes + h_nrm.nbytes + h_kp.nbytes), "d2h_bytes_per_step": int(rows.nbytes // 2),
                "transport": "float32 rows cross PCIe by blocks of keypoints while the next block is computed; host threads widen them exactly into the float64 result",
                "api": "compute_fpfh_descriptor(keypoints_indices, cloud_points, normals, radius, n_bins=11, decorrelated=True) -> float64 (N,33)"},
        "roofline": {"kernel": dominant, "bound": "hbm", "achieved": alg[dominant] / (stages[dominant] * 1e-3) / 1e9,
                     "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": alg[dominant] / (stages[dominant] * 1e-3) / 1e9 / pk["hbm_gbs"],
                     "traffic": traffic, "algorithmic_bytes": alg[dominant],
                     "dram_frac": (traffic / (stages[dominant] * 1e-3) / 1e9 / pk["hbm_gbs"]) if traffic else None,
                     # the roof that bounds this stage: the SMs' L1 data pipes (148 x 128 B per clock at 1965 MHz);
                     # ncu (profiles/r01_fpfh_fused_metrics.txt): L1 data pipe 65 % busy — wavefronts, not bytes, fill it
                     "l1": {"peak_GBps": 148 * 128 * 1.965, "achieved_GBps": alg[dominant] / (stages[dominant] * 1e-3) / 1e9,
                            "frac": alg[dominant] / (stages[dominant] * 1e-3) / 1e9 / (148 * 128 * 1.965),
                            "ncu_l1_data_pipe_busy": 0.65},
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


def _real_shot_rows(q: int):
    """Two sets of ~q real SHOT rows (scan / rigidly moved copy of a 1M-point pair, every fifth point a query) as float64
    device tensors — what SURVEY.md 8d asks C4 to be run on: 86 % zeros, clustered near neighbours."""
    import torch

    from shot_fpfh_b200 import ops, synthetic
    from shot_fpfh_b200.device import Grid, upload

    scan, normals = synthetic.bumpy_sphere(N_POINTS, seed=0)
    ref, ref_normals, _, _, _ = synthetic.rigid_pair(scan, normals)
    radius = RADIUS_IN_SPACINGS * synthetic.mean_spacing(N_POINTS)
    stride = max(1, N_POINTS // q)
    rows = []
    grid = Grid()
    for pts, nrm in ((scan, normals), (ref, ref_normals)):
        p_dev, n_dev = upload(pts), upload(nrm)
        grid.build(p_dev, n_dev, radius)
        kp = p_dev[::stride][:q].contiguous()
        d, _, _ = ops.shot_single_scale(grid, kp, radius, MIN_NB, True, out_dtype=torch.float32)
        rows.append(d.double())
    grid.close()
    return rows[0], rows[1]


def bench_match(args, pk, q: int = 200_000):
    """C4: 200k x 200k 352-d exact NN (+ second NN for the ratio test): shortlist GEMM on tensor cores, float64 re-rank,
    certificate (+ exhaustive redo of what it rejects) — on synthetic sparse unit rows and on real SHOT rows."""
    import torch

    from shot_fpfh_b200 import ops, synthetic
    from shot_fpfh_b200.matching import matching as mm

    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    steps = max(3, args.steps // 3)
    flops = 2.0 * q * q * 352

    def device_leg(a, b):
        qa = int(a.shape[0])
        state = {}

        def step(stage):
            with mark(stage, "nonempty_pack"):
                ra, ta = ops.nonempty_rows(a, want_absmax=True)
                rb, tb = ops.nonempty_rows(b, want_absmax=True)
                scale = mm.pack_scale(mm.largest(ta, tb))
                ap, an = ops.match_pack(a, ra, scale)
                bp, bn = ops.match_pack(b, rb, scale)
                bmax = float(bn.max().sqrt().item())
            with mark(stage, "shortlist_gemm"):
                score, cand = ops.match_topk(ap, bp, bn, 8, 0, True)
            with mark(stage, "rerank_f64"):
                nn, d1, d2 = ops.match_rerank(a, ra, b, rb, cand)
            with mark(stage, "certificate"):
                flags = ops.match_certify(score, an, d1, d2, scale, bmax, 352, int(rb.shape[0]), True)
                which = torch.nonzero(flags).squeeze(1)
                state["fallback_rows"] = int(which.shape[0])
                if state["fallback_rows"]:
                    norm_bound = 1.01 * (float(an.max().sqrt().item()) + bmax) + 352 ** 0.5 * 2.0 ** -23
                    mm.exhaustive_redo(a, ra, which, b, rb, d2[which], scale, norm_bound)

        ms, stages = timed_steps(step, steps, min(args.warmup, 3), lambda: flush_buf.fill_(1))
        tf = 2.0 * qa * int(b.shape[0]) * 352 / (stages["shortlist_gemm"] * 1e-3) / 1e12
        return {"value": qa / (ms * 1e-3), "unit": "match queries/s", "ms_per_step": ms, "steps": steps,
                "fallback_rows": state["fallback_rows"],
                "roofline": {"kernel": "topk_tc_kernel", "bound": "tensor", "achieved": tf, "peak": pk["tflops"],
                             "unit": "TFLOP/s", "frac": tf / pk["tflops"], "flops": 2.0 * qa * int(b.shape[0]) * 352,
                             "per_stage_ms": stages}}

    a = torch.from_numpy(synthetic.sparse_unit_rows(q, 352, seed=2)).cuda().double()
    b = torch.from_numpy(synthetic.sparse_unit_rows(q, 352, seed=3)).cuda().double()
    out = device_leg(a, b)
    out["workload"] = (f"C4: {q} x {q} x 352 exact nearest + second-nearest neighbour (certified), synthetic sparse unit rows, "
                       "1 GPU")
    # end to end through the reference-shaped call: two (q, 352) float64 host arrays in (pinned), index pairs out
    from shot_fpfh_b200.matching import basic_matching

    host_a, host_b = a.cpu().numpy(), b.cpu().numpy()
    h_a, h_b = _pinned(host_a), _pinned(host_b)
    e2e_times = []
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pairs = basic_matching(h_a, h_b)
        if i >= 1:
            e2e_times.append(time.perf_counter() - t0)
    e2e_ms = float(np.mean(e2e_times)) * 1e3
    out["e2e"] = {"value": q / (e2e_ms * 1e-3), "unit": "match queries/s", "ms_per_step": e2e_ms,
                  "h2d_bytes_per_step": int(h_a.nbytes + h_b.nbytes), "d2h_bytes_per_step": int(pairs[0].nbytes + pairs[1].nbytes),
                  "fallback_rows": mm.LAST_STATS["fallback_rows"], "api": "basic_matching(scan_descriptors, ref_descriptors)"}
    try:
        out["cpu_baseline"] = cpu_match_sample(host_a, host_b, q, min(args.cpu_seconds, 10.0))
    except Exception as exc:  # noqa: BLE001
        out["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"}
    del a, b, h_a, h_b, host_a, host_b
    # ---- the same on real SHOT rows, and through the API with the descriptors handed over on the device ----
    try:
        ra, rb = _real_shot_rows(q)
        real = device_leg(ra, rb)
        real["workload"] = (f"C4 on real rows: {int(ra.shape[0])} x {int(rb.shape[0])} SHOT descriptors of a 1M-point rigid pair "
                            "(every fifth point a query)")
        del ra, rb
        from shot_fpfh_b200.descriptors import ShotMultiprocessor

        scan, normals = synthetic.bumpy_sphere(N_POINTS, seed=0)
        ref, ref_normals, _, _, _ = synthetic.rigid_pair(scan, normals)
        radius = RADIUS_IN_SPACINGS * synthetic.mean_spacing(N_POINTS)
        stride = max(1, N_POINTS // q)
        with ShotMultiprocessor(min_neighborhood_size=MIN_NB, verbose=False) as shot:
            d_scan = shot.compute_descriptor_single_scale(scan, normals, scan[::stride][:q], radius)
            d_ref = shot.compute_descriptor_single_scale(ref, ref_normals, ref[::stride][:q], radius)
        times = []
        for i in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pairs = basic_matching(d_scan, d_ref)
            if i >= 1:
                times.append(time.perf_counter() - t0)
        ms_api = float(np.mean(times)) * 1e3
        real["e2e_from_descriptors"] = {
            "value": d_scan.shape[0] / (ms_api * 1e-3), "unit": "match queries/s", "ms_per_step": ms_api,
            "handoff": mm.LAST_STATS["handoff"], "fallback_rows": mm.LAST_STATS["fallback_rows"], "matches": int(pairs[0].shape[0]),
            "what": "basic_matching on the very arrays compute_descriptor_single_scale returned (pipeline.py:376-399): the "
                    "float32 rows they left on the device are matched, nothing crosses PCIe but the index pairs"}
        out["real_shot_rows"] = real
    except Exception as exc:  # noqa: BLE001
        out["real_shot_rows"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def bench_distributed(args, dist, rank, world):
    """
    N > 1 only: the two sharded paths that DO exchange data, strong scaling of one job over the ranks,
    through `shot_fpfh_b200.distributed` with host arrays in: FPFH C3 (SPFH by blocks of the cell-sorted cloud, ONE
    all-gather of the SPFH rows over NVLink, FPFH by keypoint blocks) and matching C4 (target set sharded, every rank
    emits its exact nearest / second nearest, ONE all-gather of a (Q, 3) float64 tensor, merge). Wall clock between
    barriers, max over ranks, best of 3; the stages of the best run by CUDA events (max over ranks), the collective's
    share separated from the uploads. Results are checked across ranks (same checksum everywhere).
    """
    import torch

    from shot_fpfh_b200 import distributed, synthetic

    out = {}
    pts, normals = synthetic.bumpy_sphere(N_POINTS, seed=0)
    radius = RADIUS_IN_SPACINGS * synthetic.mean_spacing(N_POINTS)
    h_pts, h_nrm = _pinned(pts), _pinned(normals)
    kp = np.arange(N_POINTS, dtype=np.int64)

    def timed(fn, reps=3):
        best, best_stages = float("inf"), {}
        for _ in range(reps + 1):  # first call warms the pools
            timings = {"enabled": True}
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            fn(timings)
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            names = sorted(timings.get("ms", {}))
            stage = torch.tensor([timings["ms"][k] for k in names], device="cuda", dtype=torch.float64)
            dist.all_reduce(stage, op=dist.ReduceOp.MAX)
            if float(t.item()) < best:
                best, best_stages = float(t.item()), dict(zip(names, (float(x) for x in stage.tolist())))
        return best * 1e3, best_stages

    state = {}

    def run_fpfh(timings):
        state["fpfh"] = distributed.fpfh(kp, h_pts, h_nrm, radius, 11, decorrelated=True, gather=False, timings=timings)

    ms, stages = timed(run_fpfh)
    mine, rows = state["fpfh"]
    check = torch.stack([rows.double().sum(), torch.tensor(float(mine.shape[0]), device="cuda", dtype=torch.float64)])
    dist.all_reduce(check)
    out["fpfh_c3_sharded"] = {
        "workload": f"C3 over {world} GPUs: every rank gets the cloud (an N-th over PCIe + one all-gather), builds the grid, "
                    "SPFH of its block of the cell-sorted cloud, one all-gather of SPFH rows (144 MB in all), FPFH of the "
                    "keypoints of its block; rows left on the devices",
        "ms": ms, "value": N_POINTS / (ms * 1e-3), "unit": "descriptors/s", "scaling": "strong", "stages_ms": stages,
        "collective": {"what": "all_gather_into_tensor of the SPFH rows", "bytes_total": int(N_POINTS * 36 * 4),
                       "ms": stages.get("spfh_all_gather")},
        "rows_over_all_ranks": int(check[1].item()), "checksum": float(check[0].item()),
    }
    del state["fpfh"], rows, mine
    qm = 200_000
    a = _pinned(synthetic.sparse_unit_rows(qm, 352, seed=2).astype(np.float64))
    b = _pinned(synthetic.sparse_unit_rows(qm, 352, seed=3).astype(np.float64))

    def run_match(timings):
        state["match"] = distributed.nearest_neighbors(a, b, timings=timings)

    ms, stages = timed(run_match, reps=2)
    rows_a, nn, d1, d2 = state["match"]
    digest = torch.tensor([float(nn.sum()), float(d1.sum())], device="cuda", dtype=torch.float64)
    lo, hi = digest.clone(), digest.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["match_c4_sharded"] = {
        "workload": f"C4 over {world} GPUs: {qm} x {qm} x 352, target set sharded; a rank uploads its block of both sets (the "
                    "scan rows are completed over NVLink), searches its shard exactly (certified), one all-gather of a "
                    "(Q, 3) float64 tensor, merge",
        "ms": ms, "value": qm / (ms * 1e-3), "unit": "match queries/s", "scaling": "strong", "stages_ms": stages,
        "collective": {"what": "all_gather_into_tensor of (nearest, d1, d2) per rank + merge", "bytes_total": int(qm * 24 * world),
                       "ms": stages.get("gathered_and_merged")},
        "same_result_on_every_rank": bool(torch.equal(lo, hi)), "nn_checksum": float(digest[0].item()),
    }
    del state["match"]

    # ---- SHOT over the ranks, one cloud: query blocks on a replicated grid vs the halo partition -------------------
    kp_xyz = _pinned(np.ascontiguousarray(pts[::10]))

    def run_blocks(timings):
        state["shot"] = distributed.shot_single_scale(h_pts, h_nrm, kp_xyz, radius, True, MIN_NB, gather=False)

    def run_slabs(timings):
        state["shot"] = distributed.shot_single_scale(h_pts, h_nrm, kp_xyz, radius, True, MIN_NB, gather=False, partition="slabs")

    def wall(fn):
        best = float("inf")
        for _ in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            fn(None)
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best * 1e3

    out["shot_one_cloud_sharded"] = {
        "workload": f"one 1M-point cloud, {kp_xyz.shape[0]} keypoints over {world} GPUs, host arrays in, rows left on the devices",
        "query_blocks_replicated_grid_ms": wall(run_blocks), "halo_slabs_ms": wall(run_slabs), "scaling": "strong",
        "note": "blocks: every rank uploads and grid-sorts the whole cloud; slabs: an N-th of the upload + one all-gather, "
                "the grid over the rank's slab of cell layers + 2 layers of halo (bit-identical rows)",
    }
    del state["shot"]

    # ---- self-test over NCCL: the sharded drivers against the single-GPU API on the same inputs (what
    #      tests/test_gpu_distributed.py::test_two_rank_nccl checks, run here so that every scaling run carries it) ----
    from shot_fpfh_b200.descriptors import ShotMultiprocessor, compute_fpfh_descriptor
    from shot_fpfh_b200.matching import basic_matching

    n_small = 60_000
    s_pts, s_nrm = synthetic.bumpy_sphere(n_small, seed=9)
    s_radius = RADIUS_IN_SPACINGS * synthetic.mean_spacing(n_small)
    s_kp = np.arange(0, n_small, 7)
    ok = {}
    got_b = distributed.shot_single_scale(s_pts, s_nrm, s_pts[s_kp], s_radius, True, MIN_NB, gather=True).cpu().numpy()
    got_h = distributed.shot_single_scale(s_pts, s_nrm, s_pts[s_kp], s_radius, True, MIN_NB, gather=True,
                                          partition="slabs").cpu().numpy()
    got_f = distributed.fpfh(s_kp, s_pts, s_nrm, s_radius, 11, True, gather=True, out_dtype=torch.float64).cpu().numpy()
    sa = synthetic.sparse_unit_rows(6000, 352, seed=5).astype(np.float64)
    sb = synthetic.sparse_unit_rows(9001, 352, seed=6).astype(np.float64)
    sb[8000] = sb[17]
    m_rows, m_nn, m_d1, _ = distributed.nearest_neighbors(sa, sb)
    if rank == 0:
        with ShotMultiprocessor(min_neighborhood_size=MIN_NB, verbose=False) as shot:
            want = shot.compute_descriptor_single_scale(s_pts, s_nrm, s_pts[s_kp], s_radius)
        ok["shot_query_blocks_bit_identical"] = bool(np.array_equal(got_b.astype(np.float64), want))
        ok["shot_halo_slabs_bit_identical"] = bool(np.array_equal(got_h.astype(np.float64), want))
        ok["fpfh_bit_identical"] = bool(np.array_equal(got_f, compute_fpfh_descriptor(s_kp, s_pts, s_nrm, s_radius, 11, True,
                                                                                        verbose=False)))
        w_rows, w_nn = basic_matching(sa, sb)
        ok["matching_identical"] = bool(np.array_equal(m_rows, w_rows) and np.array_equal(m_nn, w_nn))
    flag = torch.tensor([float(all(ok.values())) if rank == 0 else 1.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["nccl_selftest"] = {"what": f"sharded drivers over {world} ranks (NCCL) against the single-GPU API, 60k-point cloud / "
                                    "6000 x 9001 rows", **ok, "passed": bool(flag.item() == 1.0)}
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
    if world > 1:  # the ranks of a node share its cores: a disjoint set each, for the host threads that rebuild results
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)
        except (AttributeError, OSError):
            pass
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
    if world > 1:
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
            "workload_stats": res["stats"],
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

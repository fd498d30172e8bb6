"""
C5-sized sanity run of the hot path (BASELINE.json configs[4] names a 10M-point scan pair): SHOT + FPFH-33 on a 10M-point
synthetic cloud with ~1M grid-selected queries, then matching of the two ~1M-row descriptor sets of a rigid pair.
Prints device-resident timings (CUDA events) and peak device memory. Not a parity test (see tests/ for those).

    python scripts/run_c5_scale.py [n_points]
"""

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from shot_fpfh_b200 import ops, synthetic  # noqa: E402
from shot_fpfh_b200.device import Grid, upload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
t0 = time.perf_counter()
scan, normals = synthetic.bumpy_sphere(n, seed=0)
s = synthetic.mean_spacing(n)
radius = 5.0 * s
print(f"generated {n} points in {time.perf_counter() - t0:.1f} s; radius {radius:.5f}")


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


p, nr = upload(scan), upload(normals)
kp_idx, ms = timed(lambda: ops.voxel_subsample(p, 3.75 * s))
print(f"voxel subsampling -> {kp_idx.shape[0]} queries: {ms:.2f} ms")
kp = p[kp_idx].contiguous()
grid = Grid()
_, ms_grid = timed(lambda: grid.build(p, nr, radius))
(offsets, nbr, _, _), ms_search = timed(lambda: ops.radius_csr(grid, kp, radius))
lrf, ms_lrf = timed(lambda: ops.shot_lrf(grid, kp, radius, offsets, nbr))
desc, ms_desc = timed(lambda: ops.shot_descriptor(grid, kp, radius, offsets, nbr, lrf, 10, True, out_dtype=torch.float32))
q = kp.shape[0]
total = ms_grid + ms_search + ms_lrf + ms_desc
print(f"SHOT {q} queries, {nbr.shape[0]} pairs: grid {ms_grid:.2f} + search {ms_search:.2f} + lrf {ms_lrf:.2f} + "
      f"descriptor {ms_desc:.2f} = {total:.2f} ms -> {q / total * 1e3 / 1e6:.1f} M descriptors/s")
norms = desc.norm(dim=1)
assert torch.isfinite(desc).all() and bool(((norms - 1).abs() < 1e-4).logical_or(norms == 0).all())
del offsets, nbr, lrf
(off2, nbr2, _, d2), ms_s2 = timed(lambda: ops.radius_csr(grid, None, radius, want_dist=True))
rows, ms_spfh = timed(lambda: ops.spfh(grid, off2, nbr2, 11, True))
f, ms_fpfh = timed(lambda: ops.fpfh(grid, off2, nbr2, d2, rows, kp_idx, out_dtype=torch.float32))
print(f"FPFH-33 on {q} keypoints (SPFH on all {n} points, {nbr2.shape[0]} pairs): search {ms_s2:.2f} + spfh {ms_spfh:.2f} + "
      f"fpfh {ms_fpfh:.2f} ms")
del off2, nbr2, d2, rows, f
print(f"peak device memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
a = desc[: min(q, 1_000_000)].double()
b = a[torch.randperm(a.shape[0], device=a.device)] + 1e-3 * torch.rand_like(a)


def match():
    ra, rb = ops.nonempty_rows(a), ops.nonempty_rows(b)
    ap, _ = ops.match_pack(a, ra, 1.0)
    bp, bn = ops.match_pack(b, rb, 1.0)
    _, cand = ops.match_topk(ap, bp, bn, 8, 0, True)
    return ops.match_rerank(a, ra, b, rb, cand)


(nn, d1, _), ms_match = timed(match)
print(f"matching {a.shape[0]} x {b.shape[0]} x 352: {ms_match:.1f} ms -> {a.shape[0] / ms_match * 1e3 / 1e6:.2f} M queries/s, "
      f"{2.0 * a.shape[0] * b.shape[0] * 352 / ms_match / 1e9:.0f} TFLOP/s-equivalent overall")
print(f"peak device memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")

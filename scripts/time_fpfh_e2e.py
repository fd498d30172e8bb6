"""Phase times of the end-to-end FPFH call on the C3 workload (host pinned arrays in, float64 host array out)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from shot_fpfh_b200 import synthetic  # noqa: E402
from shot_fpfh_b200.descriptors.fpfh import _cached_grid, fpfh_device  # noqa: E402
from shot_fpfh_b200.device import download, download_widened, upload  # noqa: E402

n = bench.N_POINTS
pts, normals = synthetic.bumpy_sphere(n, seed=0)
radius = bench.RADIUS_IN_SPACINGS * synthetic.mean_spacing(n)
h_pts, h_nrm = bench._pinned(pts), bench._pinned(normals)
kp = np.arange(n, dtype=np.int64)


def sync():
    torch.cuda.synchronize()


for it in range(4):
    sync()
    t = [time.perf_counter()]
    p, q = upload(h_pts), upload(h_nrm)
    k = upload(kp, torch.int64)
    sync(); t.append(time.perf_counter())
    grid = _cached_grid().build(p, q, radius)
    sync(); t.append(time.perf_counter())
    out, _ = fpfh_device(grid, k, radius, 11, True, out_dtype=torch.float32)
    sync(); t.append(time.perf_counter())
    res = download_widened(out)
    t.append(time.perf_counter())
    res64 = download(out.double())
    t.append(time.perf_counter())
    for blocks in (1, 2, 8):
        download_widened(out, blocks=blocks)
        t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("upload %.2f  grid %.2f  fpfh %.2f  download_widened %.2f  (float64 download %.2f; widened 1/2/8 blocks %.2f %.2f %.2f) ms"
          % tuple(d))

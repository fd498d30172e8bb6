"""Small driver for ncu captures of the C2 SHOT step (grid build + fused driver): python scripts/profile_c2_step.py [steps]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from shot_fpfh_b200 import ops  # noqa: E402
from shot_fpfh_b200.device import Grid, upload  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pts, normals, kp, radius = bench.make_shot_workload(0)
p_dev, n_dev, k_dev = upload(pts), upload(normals), upload(kp)
grid = Grid().set_speculative(builds=True, shot_lists=True)
out = torch.empty((kp.shape[0], 352), dtype=torch.float32, device="cuda")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for _ in range(steps):
    flush.fill_(1)
    grid.build(p_dev, n_dev, radius)
    ops.shot_single_scale(grid, k_dev, radius, bench.MIN_NB, True, out=out)
torch.cuda.synchronize()
assert grid.poll() == 0
print("ok", float(out.sum().item()))

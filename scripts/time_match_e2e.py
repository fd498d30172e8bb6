"""Timeline of the pipelined end-to-end matching call (C4 sizes): where the host blocks."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from shot_fpfh_b200 import synthetic  # noqa: E402
import shot_fpfh_b200.matching.matching as mm  # noqa: E402

q = 200_000
a = bench._pinned(synthetic.sparse_unit_rows(q, 352, seed=2).astype(np.float64))
b = bench._pinned(synthetic.sparse_unit_rows(q, 352, seed=3).astype(np.float64))
print("pinned?", torch.from_numpy(a).is_pinned(), torch.from_numpy(b[1000:5000]).is_pinned())
for rows in (mm._PIPELINE_CHUNK_ROWS, 2 * mm._PIPELINE_CHUNK_ROWS, 50_000):
    mm._PIPELINE_CHUNK_ROWS = rows
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = mm.basic_matching(a, b)
        t1 = time.perf_counter()
        print(f"scan rows per chunk {rows}: {1e3 * (t1 - t0):.2f} ms")
mm._PIPELINE_MIN_ROWS = 10**9
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = mm.basic_matching(a, b)
    print(f"one piece: {1e3 * (time.perf_counter() - t0):.2f} ms")
# raw copy timing
dev = torch.empty(a.shape, dtype=torch.float64, device="cuda")
ta = torch.from_numpy(a)
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev.copy_(ta, non_blocking=True)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"copy 563 MB: enqueue {1e3 * (t1 - t0):.2f} ms, total {1e3 * (time.perf_counter() - t0):.2f} ms")

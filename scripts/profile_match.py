"""Small driver for ncu captures of the shortlist GEMM: python scripts/profile_match.py [qa] [qb] [reps]."""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from shot_fpfh_b200 import ops, synthetic  # noqa: E402

qa = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
qb = int(sys.argv[2]) if len(sys.argv) > 2 else 51200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
a = torch.from_numpy(synthetic.sparse_unit_rows(qa, 352, seed=2)).cuda().double()
b = torch.from_numpy(synthetic.sparse_unit_rows(qb, 352, seed=3)).cuda().double()
ra, rb = ops.nonempty_rows(a), ops.nonempty_rows(b)
ap, _ = ops.match_pack(a, ra, 1.0)
bp, bn = ops.match_pack(b, rb, 1.0)
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.match_topk(ap, bp, bn, 8, 0, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"topk_tc {qa} x {qb}: {ms:.3f} ms, {2.0 * qa * qb * 352 / ms / 1e9:.1f} TFLOP/s")

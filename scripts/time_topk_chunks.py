import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shot_fpfh_b200 import ops, synthetic
q = 200_000
a = torch.from_numpy(synthetic.sparse_unit_rows(q, 352, seed=2)).cuda().double()
b = torch.from_numpy(synthetic.sparse_unit_rows(q, 352, seed=3)).cuda().double()
ra, rb = ops.nonempty_rows(a), ops.nonempty_rows(b)
ap, _ = ops.match_pack(a, ra, 1.0)
def ev(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for parts in (1, 2, 4, 8):
    n = q // parts
    bp, bn = ops.match_pack(b[:n].contiguous(), ops.nonempty_rows(b[:n].contiguous()), 1.0)
    print(parts, "topk ms per chunk", ev(lambda: ops.match_topk(ap, bp, bn, 8, 0, True)), " x parts =", parts * ev(lambda: ops.match_topk(ap, bp, bn, 8, 0, True)))
chunk = b[:50000]
print("nonempty", ev(lambda: ops.nonempty_rows(chunk)), "absmax", ev(lambda: chunk.abs().max()), "pack", ev(lambda: ops.match_pack(chunk, rb[:50000], 1.0)), "packA", ev(lambda: ops.match_pack(a, ra, 1.0)))

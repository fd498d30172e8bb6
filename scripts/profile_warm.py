import cProfile, pstats, sys, time, os
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import bench
from shot_fpfh_b200.descriptors import ShotMultiprocessor
pts, normals, kp, radius = bench.make_shot_workload(0)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = a; return t.numpy()
h=[pinned(x) for x in (pts,normals,kp)]
print("is_pinned:", torch.from_numpy(h[0]).is_pinned())
if len(sys.argv) > 1 and sys.argv[1] == "nohandoff":
    import shot_fpfh_b200.descriptors.shot_parallelization as sp
    sp.remember_device_rows = lambda *a, **k: None
with ShotMultiprocessor(min_neighborhood_size=10, verbose=False) as shot:
    for i in range(6):
        torch.cuda.synchronize(); t0=time.perf_counter()
        d = shot.compute_descriptor_single_scale(h[0],h[1],h[2],radius)
        print("warm %.2f ms"%(1e3*(time.perf_counter()-t0)))
    prof=cProfile.Profile(); prof.enable()
    d = shot.compute_descriptor_single_scale(h[0],h[1],h[2],radius)
    prof.disable(); pstats.Stats(prof).sort_stats("cumulative").print_stats(14)

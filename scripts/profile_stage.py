"""Runs one stage of bench.py a few times (for ncu launch lists): python scripts/profile_stage.py shot|fpfh|match [steps]"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

args = types.SimpleNamespace(steps=int(sys.argv[2]) if len(sys.argv) > 2 else 2, warmup=1, cpu_seconds=2.0)
pk = bench.peaks()
what = sys.argv[1]
if what == "shot":
    r = bench.bench_shot(args, None, 0, 1, pk)
    print(r["ms"], r["roofline"]["per_stage"], r["e2e"]["ms_per_step"])
elif what == "fpfh":
    print(bench.bench_fpfh(args, pk))
else:
    print(bench.bench_match(args, pk))

"""GX_FILL_STATS=1 python tools/fill_stats.py [workload] -- where do the fill kernel's warps spend their cycles?"""
import os, sys
os.environ["GX_FILL_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib
_lib.ensure_init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "corona45"
w = bench.build_workload(name, 0, 1, 1_000_000)
plan = gx.Plan(w["len1"], w["len2"], bench.SCORES, w["is_local"], traceback=w["traceback"])
plan.upload(w["blob"], w["off1"], w["off2"])
for _ in range(3):
    plan.execute()
top, bnd, tile, s1, n = [plan.stat(k) for k in range(10, 15)]
print(f"{name}: K={int(plan.stat(15))} fill {plan.fill_ms:.3f} ms walk {plan.walk_ms:.3f} ms tiles {int(n)}")
print(f"  per-warp cycles inside tiles: {tile:.3e}; top-dependency wait {100*top/tile:.1f}%  left-boundary wait {100*bnd/tile:.1f}%  s1 TMA wait {100*s1/tile:.1f}%")
print(f"  avg tile {tile/n/1.963e3:.1f} us, avg top wait {top/n/1.963e3:.1f} us, avg boundary wait {bnd/n/1.963e3:.1f} us")

"""GX_K=.. GX_CHAIN1=.. python tools/shape_stats.py m n -- wait breakdown of the fill kernel for one m x n global score-only pair"""
import os, sys
os.environ["GX_FILL_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib, workloads as wl
_lib.ensure_init(0)
m, n = int(sys.argv[1]), int(sys.argv[2])
a, b = wl.long_pair(max(m, n))
plan = gx.Plan([m], [n], wl.CONFIG_TOML, False, traceback=False)
plan.upload(np.concatenate([a[:m], b[:n]]), [0], [m])
for _ in range(2):
    plan.execute()
top, bnd, tile, s1, nt = [plan.stat(k) for k in range(10, 15)]
K = int(plan.stat(15)); S = -(-n // (32 * K)); rows = min(m, 4096)
print(f"{m}x{n} K={K} chain1={int(plan.stat(17))} strips={S}: fill {plan.fill_ms:.3f} ms = {m*n/plan.fill_ms/1e6:.0f} GCUPS; "
      f"critical path {plan.fill_ms*1e-3*1.965e9/(m+31):.0f} clk/row")
print(f"   in tiles: top wait {100*top/tile:.1f}%  boundary wait {100*bnd/tile:.1f}%  compute {(tile-top-bnd-s1)/nt/rows:.0f} clk/step; "
      f"lag per strip = {(plan.fill_ms*1e-3*1.965e9 - (m+31)*(tile-top-bnd-s1)/nt/rows)/max(S-1,1):.0f} clk")

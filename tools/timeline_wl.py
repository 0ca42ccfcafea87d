"""GX_FILL_STATS=2 python tools/timeline_wl.py workload rank world -- per-tile schedule summary for one rank's shard of a bench workload"""
import os, sys
os.environ["GX_FILL_STATS"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib
lib = _lib.ensure_init(0)
name, rank, world = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
w = bench.build_workload(name, rank, world, 1_000_000)
plan = gx.Plan(w["len1"], w["len2"], bench.SCORES, w["is_local"], traceback=w["traceback"])
plan.upload(w["blob"], w["off1"], w["off2"])
for _ in range(3):
    plan.execute()
nt = int(plan.stat(8))
tl = np.zeros(nt * 4, np.uint64)
_lib.check(lib.gx_plan_debug_timeline(plan._h, tl.ctypes.data, tl.size))
tl = tl.reshape(nt, 4).astype(np.int64)
t0 = tl[:, 0].min()
pair = (tl[:, 3] >> 48) & 0xffff; p = (tl[:, 3] >> 32) & 0xffff; s = (tl[:, 3] >> 12) & 0xfffff; sm = tl[:, 3] & 0xfff
K = int(plan.stat(15))
print(f"{name} rank {rank}/{world}: pairs {len(w['len1'])} K={K} chain1={int(plan.stat(17))} fill {plan.fill_ms:.3f} ms walk {plan.walk_ms:.3f} ms tiles {nt}")
S = int(s.max()) + 1
for q in range(min(2, len(w["len1"]))):
    for pp in range(int(p[pair == q].max()) + 1):
        sel = (pair == q) & (p == pp)
        o = np.argsort(s[sel])
        dp0 = (tl[sel, 1][o] - t0) / 1e3; end = (tl[sel, 2][o] - t0) / 1e3
        print(f"  pair {q} panel {pp}: strip0 dp0 {dp0[0]:8.1f} end {end[0]:8.1f} | last strip dp0 {dp0[-1]:8.1f} end {end[-1]:8.1f} | median run {np.median(end-dp0):7.1f} us, median lag {np.median(np.diff(dp0)):5.2f} us")
top, bnd, tile, s1w, ntl = [plan.stat(k) for k in range(10, 15)]
print(f"  in tiles: top wait {100*top/tile:.1f}%  boundary wait {100*bnd/tile:.1f}%  compute {(tile-top-bnd-s1w)/ntl/4096:.0f} clk/step (4096-row tiles)")
busy = ((tl[:, 2] - tl[:, 1]).sum()) / 1e3
print(f"  sum of tile run times {busy/1e3:.2f} ms over {nt} tiles; kernel {plan.fill_ms:.2f} ms; distinct SMs used {len(set(sm.tolist()))}")

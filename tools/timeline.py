"""GX_FILL_STATS=2 python tools/timeline.py m n -- per-tile schedule of the fill kernel for one m x n pair"""
import os, sys
os.environ["GX_FILL_STATS"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib, workloads as wl
lib = _lib.ensure_init(0)
m, n = int(sys.argv[1]), int(sys.argv[2])
a, b = wl.long_pair(max(m, n))
TB = os.environ.get("TB", "0") == "1"
LOCAL = os.environ.get("LOCAL", "0") == "1"
plan = gx.Plan([m], [n], wl.CONFIG_TOML, LOCAL, traceback=TB)
plan.upload(np.concatenate([a[:m], b[:n]]), [0], [m])
for _ in range(2):
    plan.execute()
nt = int(plan.stat(8))
tl = np.zeros(nt * 4, np.uint64)
_lib.check(lib.gx_plan_debug_timeline(plan._h, tl.ctypes.data, tl.size))
tl = tl.reshape(nt, 4)
t0 = tl[:, 0].min()
p = (tl[:, 3] >> np.uint64(32)) & np.uint64(0xffff); s = (tl[:, 3] >> np.uint64(12)) & np.uint64(0xfffff); sm = tl[:, 3] & np.uint64(0xfff)
print(f"{m}x{n} K={int(plan.stat(15))} fill {plan.fill_ms:.3f} ms tiles {nt}")
order = np.lexsort((s, p))
for k in order[: int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"p={int(p[k])} s={int(s[k]):4d} sm={int(sm[k]):3d} take {(tl[k,0]-t0)/1e3:9.1f} us  dp0 {(tl[k,1]-t0)/1e3:9.1f}  end {(tl[k,2]-t0)/1e3:9.1f}  run {(tl[k,2]-tl[k,1])/1e3:8.1f} us")
d = np.diff(np.array([tl[k, 1] for k in order if p[k] == 0], dtype=np.float64))
print("panel 0: start-to-start lag between adjacent strips: median %.2f us, mean %.2f us, max %.2f" % (np.median(d) / 1e3, d.mean() / 1e3, d.max() / 1e3))
rows0 = min(m, 4096)
run0 = np.array([float(tl[k, 2] - tl[k, 1]) for k in order if p[k] == 0])
step_ns = np.median(run0) / (rows0 + 31)
print("panel 0: tile run time median %.1f us = %.1f ns per step (%d rows + 31 steps of skew): lag = %.1f steps per strip; %d strips"
      % (np.median(run0) / 1e3, step_ns, rows0, np.median(d) / step_ns, len(run0)))

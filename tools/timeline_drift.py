"""GX_FILL_STATS=2 python tools/timeline_drift.py m n -- how the start times of adjacent strips drift apart panel by panel"""
import os, sys
os.environ["GX_FILL_STATS"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib, workloads as wl
lib = _lib.ensure_init(0)
m, n = int(sys.argv[1]), int(sys.argv[2])
a, b = wl.long_pair(max(m, n))
plan = gx.Plan([m], [n], wl.CONFIG_TOML, os.environ.get("LOCAL", "0") == "1", traceback=os.environ.get("TB", "0") == "1")
plan.upload(np.concatenate([a[:m], b[:n]]), [0], [m])
for _ in range(2):
    plan.execute()
nt = int(plan.stat(8))
tl = np.zeros(nt * 4, np.uint64)
_lib.check(lib.gx_plan_debug_timeline(plan._h, tl.ctypes.data, tl.size))
tl = tl.reshape(nt, 4).astype(np.int64)
t0 = tl[:, 0].min()
p = (tl[:, 3] >> 32) & 0xffff; s = (tl[:, 3] >> 12) & 0xfffff
P, S = int(p.max()) + 1, int(s.max()) + 1
dp0 = np.zeros((P, S)); end = np.zeros((P, S)); take = np.zeros((P, S))
dp0[p, s] = (tl[:, 1] - t0) / 1e3; end[p, s] = (tl[:, 2] - t0) / 1e3; take[p, s] = (tl[:, 0] - t0) / 1e3
print(f"{m}x{n} K={int(plan.stat(15))} chain1={int(plan.stat(17))} fill {plan.fill_ms:.3f} ms; panels {P} strips {S}")
for pp in sorted(set([0, 1, 2, P // 4, P // 2, P - 2, P - 1])):
    if pp < 0 or pp >= P: continue
    lag = np.diff(dp0[pp])
    run = end[pp] - dp0[pp]
    gap = dp0[pp] - (end[pp - 1] if pp else 0)      # previous panel of the same strip finished -> this one starts computing
    print(f"panel {pp:4d}: strip0 dp0 {dp0[pp,0]:10.1f} us; lag/strip median {np.median(lag):6.2f} mean {lag.mean():6.2f} us; "
          f"tile run median {np.median(run):7.1f} us; panel hand-over gap median {np.median(gap):6.2f} us; last strip end {end[pp,-1]:10.1f}")

"""python tools/strip_probe.py -- lone-warp speed of the fill kernel: m rows x (S strips of 32*K columns), global score only"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib, workloads as wl
_lib.ensure_init(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
a, b = wl.long_pair(max(m, 1 << 16))
K = int(os.environ.get("GX_K", "16"))
for strips in [1, 2, 4, 16, 64]:
    n = strips * 32 * K
    plan = gx.Plan([m], [n], wl.CONFIG_TOML, False, traceback=False)
    plan.upload(np.concatenate([a[:m], b[:n]]), [0], [m])
    for _ in range(2):
        plan.execute()
    ms = plan.fill_ms
    steps = m + 31 + 0.0
    print(f"K={K} chain1={int(plan.stat(17))} strips={strips:3d}: fill {ms:8.3f} ms  -> {ms*1e-3*1.965e9/steps:7.1f} clk per step (first strip), {m*n/ms/1e6:8.1f} GCUPS")
    plan.close()

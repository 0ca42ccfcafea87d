"""python tools/one_strip.py [m] [strips] -- one m x (strips*32*K) global score-only pair (profiling target)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib, workloads as wl
_lib.ensure_init(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
strips = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K = int(os.environ.get("GX_K", "16"))
n = strips * 32 * K
a, b = wl.long_pair(max(m, n))
TB = os.environ.get("TB", "0") == "1"
LOCAL = os.environ.get("LOCAL", "0") == "1"
plan = gx.Plan([m], [n], wl.CONFIG_TOML, LOCAL, traceback=TB)
plan.upload(np.concatenate([a[:m], b[:n]]), [0], [m])
for _ in range(2):
    plan.execute()
print(f"TB={int(TB)} LOCAL={int(LOCAL)} K={K} chain1={int(plan.stat(17))} m={m} strips={strips}: fill {plan.fill_ms:.3f} ms -> {plan.fill_ms*1e-3*1.965e9/(m+31):.1f} clk/step")

"""python tools/shape_probe.py m n -- fill time of one m x n global score-only pair for every (K, CHAIN1)"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 3:
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
    ms = plan.fill_ms
    print(f"{m}x{n} K={int(plan.stat(15)):2d} chain1={int(plan.stat(17))}: fill {ms:9.3f} ms  {m*n/ms/1e6:8.1f} GCUPS  strips {-(-n//(32*int(plan.stat(15))))}")
else:
    for k in (4, 8, 16):
        for c in (0, 1):
            env = dict(os.environ, GX_K=str(k), GX_CHAIN1=str(c))
            subprocess.run([sys.executable, __file__, sys.argv[1], sys.argv[2], "x"], env=env)

#!/usr/bin/env python
"""fill time of one pair in every kernel mode (global/local x score/start-cell/traceback): python tools/mode_probe.py [brca2|corona1]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib, workloads as wl
_lib.ensure_init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "brca2"
if name == "brca2":
    a, b = wl.brca2_pair()
else:
    seqs, jobs = wl.corona_pairs(); a, b = seqs[0], seqs[1]
blob, off1, len1, off2, len2 = gx.pack_pairs([(a, b)])
for is_local in (False, True):
    for tb, sc in ((False, False), (False, True), (True, False)):
        plan = gx.Plan(len1, len2, wl.CONFIG_TOML, is_local, traceback=tb, start_cell=sc)
        plan.upload(blob, off1, off2)
        for _ in range(3): plan.execute()
        t = []
        for _ in range(5):
            plan.execute(); t.append(plan.fill_ms)
        print(f"{name} local={int(is_local)} traceback={int(tb)} start_cell={int(sc)} K={int(plan.stat(15))} c1={int(plan.stat(17))} fill {np.median(t):.3f} ms", flush=True)
        plan.close()

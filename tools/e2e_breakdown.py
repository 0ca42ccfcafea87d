"""Where does the end-to-end time of one gx_align_batch-equivalent call go? (create / upload / execute / fetch / destroy)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib
_lib.ensure_init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "corona45"
w = bench.build_workload(name, 0, 1, 1_000_000)
acc = {}
for it in range(6):
    t = [time.perf_counter()]
    plan = gx.Plan(w["len1"], w["len2"], bench.SCORES, w["is_local"], traceback=w["traceback"]); t.append(time.perf_counter())
    plan.upload(w["blob"], w["off1"], w["off2"]); t.append(time.perf_counter())
    plan.execute(); t.append(time.perf_counter())
    if w["traceback"]:
        plan.fetch()
    else:
        plan.fetch_scores()
    t.append(time.perf_counter())
    fill, walk = plan.fill_ms, plan.walk_ms
    plan.close(); t.append(time.perf_counter())
    if it >= 2:
        for k, nm in enumerate(["create", "upload", "execute", "fetch", "destroy"]):
            acc.setdefault(nm, []).append((t[k + 1] - t[k]) * 1e3)
        acc.setdefault("fill_dev", []).append(fill); acc.setdefault("walk_dev", []).append(walk)
print(name, {k: round(float(np.mean(v)), 3) for k, v in acc.items()}, "total", round(sum(np.mean(v) for k, v in acc.items() if not k.endswith("_dev")), 3))

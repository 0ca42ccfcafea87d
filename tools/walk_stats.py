import os, sys
os.environ["GX_WALK_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib
_lib.ensure_init(0)
w = bench.build_workload("corona45", 0, 1, 0)
plan = gx.Plan(w["len1"], w["len2"], bench.SCORES, False, traceback=True)
plan.upload(w["blob"], w["off1"], w["off2"]); plan.execute(); plan.execute()
res, ops, off = plan.fetch()
print("walk_ms", plan.walk_ms)
for q in np.argsort(-res["fill_ms"])[:6].tolist() + np.argsort(res["fill_ms"])[:2].tolist():
    st = int(res["lcs_at_first_max"][q])
    it = st & 0xffffffff; rl = (st >> 32) & 0xffff; miss = st >> 48; ringwait = int(res["start_i"][q]) >> 32
    cyc = res["fill_ms"][q]
    rc = res["walk_ms"][q]
    print(f"pair {q}: window changes {rl} (misses {miss}, {rc/max(rl,1):.0f} clk each), ring-full waits {ringwait}, walk-only {(cyc-rc)/max(it,1):.0f} cyc/iter; ops {int(res['n_ops'][q])} opens {int(res['opening_gaps'][q])} iters {it} reloads {rl} cycles {cyc:.0f} = {cyc/1.963e6:.3f} ms, {cyc/max(it,1):.0f} cyc/iter")

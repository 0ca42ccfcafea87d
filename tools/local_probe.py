import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib, workloads as wl
_lib.ensure_init(0)
a, b = wl.brca2_pair()
blob, off1, len1, off2, len2 = gx.pack_pairs([(a, b)])
plan = gx.Plan(len1, len2, wl.CONFIG_TOML, True, traceback=True)
plan.upload(blob, off1, off2)
for _ in range(4): plan.execute()
print("fill", plan.fill_ms)

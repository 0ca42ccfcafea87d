#!/usr/bin/env python
"""one-screen summary of a bench.py JSON line (last line of the file given)"""
import json
import sys

try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
except Exception as e:
    raise SystemExit(f"{sys.argv[1]}: bench line unreadable ({e})")


def row(name, c):
    r, e = c.get("roofline") or {}, c.get("e2e") or {}
    print(f"{name:13s} n={d.get('n_gpus')} value {c.get('value', 0):9.1f} ms {c.get('ms_per_step', 0):8.3f} fill {c.get('fill_ms_per_step', 0):8.3f} "
          f"walk {c.get('walk_ms_per_step', 0) or 0:6.3f} e2e {e.get('value', 0):9.1f} ({e.get('ms_per_step', 0):8.3f} ms) parity {c.get('parity_ok')} "
          f"alu_frac {r.get('frac', 0):.3f} mix_frac {r.get('frac_of_mix_ceiling', 0) or 0:.3f} cells/clk/SM {r.get('cells_per_clk_per_sm', 0):.2f} "
          f"K {c.get('K', (c.get('plan') or {}).get('K'))} {c.get('error', '')}")


row(d["config"]["workload"], d)
for k, c in (d.get("configs") or {}).items():
    row(k, c)
if d.get("clocks"):
    print("clocks", d["clocks"], "all_parity_ok", d.get("all_parity_ok"))

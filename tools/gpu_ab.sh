#!/bin/bash
# A/B of the two recurrence forms (GX_CHAIN1=0/1) on every wavefront workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline --no-k0"
for c in 0 1; do
  GX_CHAIN1=$c $B --workload corona45 --steps 10 --warmup 3 > gpurun_out/ab_corona45_c$c.json 2>&1
  GX_CHAIN1=$c $B --workload brca2_global --steps 20 --warmup 3 > gpurun_out/ab_brca2g_c$c.json 2>&1
  GX_CHAIN1=$c $B --workload brca2_local --steps 20 --warmup 3 > gpurun_out/ab_brca2l_c$c.json 2>&1
  for k in 8 16; do
    GX_CHAIN1=$c GX_K=$k $B --workload nw1m --steps 2 --warmup 1 > gpurun_out/ab_nw1m_k${k}_c$c.json 2>&1
    GX_CHAIN1=$c GX_K=$k $B --workload nw1m --length 200000 --steps 3 --warmup 1 > gpurun_out/ab_nw200k_k${k}_c$c.json 2>&1
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "GCUPS %.1f  ms %.3f fill %.3f walk %s e2e %.1f" % (d["value"], d["ms_per_step"], d["fill_ms_per_step"], d.get("walk_ms_per_step"), d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-300:])
PY

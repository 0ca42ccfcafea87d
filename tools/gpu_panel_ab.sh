#!/bin/bash
# A/B of the panel height (4096 / 2048 / 1024 rows per tile): finer tiles = tighter code band, more per-tile overhead
mkdir -p gpurun_out
for tag in "" p11 p10; do
  lib=genomics_rs_b200/libgxalign${tag:+_$tag}.so
  echo "== ${tag:-p12}"
  GX_LIB_PATH=$PWD/$lib timeout 300 python tools/sanitize_small.py 2>&1 | tail -1
  out=gpurun_out/sweep_panel_${tag:-p12}.jsonl; rm -f $out
  GX_LIB_PATH=$PWD/$lib timeout 600 python tools/sweep_kr.py --workloads brca2_global,brca2_local,corona6,corona23,corona45,nw1m --combos 8x1 --chain 0 --steps 4 --out $out > /dev/null 2>&1
  python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    r = json.loads(l)
    if 'error' in r: print(r); continue
    if r['forced'] and r['workload'] not in ('corona23',): continue
    print(f"{r['workload']:13s} K={r['K']:2d} B={r.get('batch')} res={r.get('resident')} c1={r['chain1']} forced={int(r['forced'])} fill {r['fill_ms']:9.3f} walk {r['walk_ms']:6.3f} gcups {r['gcups_fill']:8.1f} agree={r['scores_agree']}")
PY
done

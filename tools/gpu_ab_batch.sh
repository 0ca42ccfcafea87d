#!/bin/bash
# A/B of the hand-off batch length (GX_BMIN = 8 / 16 / 32 steps): three builds of the library, same sweep
mkdir -p gpurun_out
for tag in "" b16 b32; do
  lib=genomics_rs_b200/libgxalign${tag:+_$tag}.so
  [ -f $lib ] || { echo "missing $lib"; continue; }
  out=gpurun_out/sweep_batch_${tag:-b8}.jsonl
  rm -f $out
  GX_LIB_PATH=$PWD/$lib timeout 900 python tools/sweep_kr.py --workloads brca2_global,corona6,corona45,nw200k,nw1m --chain 0 --steps 4 --out $out > gpurun_out/sweep_batch_${tag:-b8}.log 2>&1
  echo "== ${tag:-b8}"; python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    r = json.loads(l)
    if 'error' in r: print(r); continue
    print(f"{r['workload']:13s} K={r['K']:2d} c1={r['chain1']} forced={int(r['forced'])} fill {r['fill_ms']:9.3f} walk {r['walk_ms']:6.3f} gcups {r['gcups_fill']:8.1f} agree={r['scores_agree']}")
PY
done

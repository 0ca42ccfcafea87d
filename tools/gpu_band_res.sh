#!/bin/bash
for br in 0 1; do
  export GX_BAND_RESIDENT=$br
  echo "== GX_BAND_RESIDENT=$br"
  python tools/mode_probe.py brca2 | grep "local=0 traceback=1"
  python tools/mode_probe.py corona1 | grep "local=0 traceback=1"
  rm -f gpurun_out/sweep_band.jsonl
  python tools/sweep_kr.py --workloads corona6,corona45 --combos 4x1 --chain 0 --steps 4 --out gpurun_out/sweep_band.jsonl > /dev/null 2>&1
  python - <<'PY'
import json
for l in open('gpurun_out/sweep_band.jsonl'):
    r = json.loads(l)
    if 'error' in r: print(r); continue
    print(f"{r['workload']:13s} K={r['K']:2d} B={r.get('batch')} res={r.get('resident')} c1={r['chain1']} forced={int(r['forced'])} fill {r['fill_ms']:9.3f} walk {r['walk_ms']:6.3f} gcups {r['gcups_fill']:8.1f}")
PY
done

#!/bin/bash
# round 2, call B: GPU tests (K = 2 tile, config-4 parity sets), the new all-configs bench line, K sweep incl. K = 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err
tail -c 1500 gpurun_out/bench_all_n1.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_all_n1.json').read().strip().splitlines()[-1])
    print('headline', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d.get('parity_ok'), 'roof', d['roofline']['frac'], d['roofline'].get('frac_of_mix_ceiling'))
    for k, c in d.get('configs', {}).items():
        print(k, c.get('value'), c.get('ms_per_step'), 'e2e', (c.get('e2e') or {}).get('value'), 'parity', c.get('parity_ok'), 'roof', (c.get('roofline') or {}).get('frac'), c.get('error'))
    print('k0', d.get('k0'))
except Exception as e:
    print('bench line unreadable', e)
PY
rm -f gpurun_out/sweep_kr.jsonl
timeout 600 python tools/sweep_kr.py --workloads brca2_global,brca2_local,corona1,corona6 --steps 5 > gpurun_out/sweep_kr.log 2>&1
tail -3 gpurun_out/sweep_kr.log

#!/bin/bash
# round 2, call F: tests with the run-time batch length + two-warp walk, all-configs bench, shard sweeps for the K/batch heuristics
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err
tail -c 600 gpurun_out/bench_all_n1.err; python tools/show_bench.py gpurun_out/bench_all_n1.json
rm -f gpurun_out/sweep_shards.jsonl
timeout 900 python tools/sweep_kr.py --workloads corona6,corona11,corona23,corona45 --combos 4x1,8x1 --batch 8,16,32 --chain 0 --steps 4 --out gpurun_out/sweep_shards.jsonl > gpurun_out/sweep_shards.log 2>&1
timeout 600 python tools/sweep_kr.py --workloads brca2_global,brca2_local,corona1,nw200k --combos 4x1,8x1 --batch 8,16,32 --chain 1 --steps 4 --out gpurun_out/sweep_shards.jsonl >> gpurun_out/sweep_shards.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_shards.jsonl'):
    r = json.loads(l)
    if 'error' in r: print(r); continue
    print(f"{r['workload']:13s} K={r['K']:2d} B={r['batch']:2d} res={r['resident']} c1={r['chain1']} forced={int(r['forced'])} fill {r['fill_ms']:9.3f} walk {r['walk_ms']:6.3f} gcups {r['gcups_fill']:8.1f} agree={r['scores_agree']}")
PY

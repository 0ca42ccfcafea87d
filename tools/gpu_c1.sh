#!/bin/bash
for c in 0 1; do
  echo "== GX_CHAIN1=$c"
  GX_CHAIN1=$c timeout 300 python bench.py --workload nw1m --steps 3 --warmup 3 --no-k0 > gpurun_out/bench_nw1m_c$c.json 2>&1; python tools/show_bench.py gpurun_out/bench_nw1m_c$c.json | head -1
  GX_CHAIN1=$c timeout 300 python bench.py --workload corona45 --steps 5 --warmup 3 --no-cpu-baseline --no-k0 > gpurun_out/bench_c45_c$c.json 2>&1; python tools/show_bench.py gpurun_out/bench_c45_c$c.json | head -1
done

#!/bin/bash
# poll-nap experiment in ticket mode + K=8 vs K=16 with 32-step batches and the IDP.4A path
mkdir -p gpurun_out
for nap in 0 100 300 1000; do
  for k in 8 16; do
    echo "== GX_POLL_NAP=$nap GX_K=$k"
    GX_POLL_NAP=$nap GX_K=$k timeout 300 python bench.py --workload corona45 --steps 5 --warmup 3 --no-cpu-baseline --no-k0 > gpurun_out/bench_c45_nap${nap}_k$k.json 2>&1; python tools/show_bench.py gpurun_out/bench_c45_nap${nap}_k$k.json | head -1
    GX_POLL_NAP=$nap GX_K=$k timeout 300 python bench.py --workload nw1m --steps 3 --warmup 3 --no-k0 > gpurun_out/bench_nw1m_nap${nap}_k$k.json 2>&1; python tools/show_bench.py gpurun_out/bench_nw1m_nap${nap}_k$k.json | head -1
  done
done

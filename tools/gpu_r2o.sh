#!/bin/bash
# A/B: score-only fill with 3 CTAs (24 warps) per SM vs 2; e2e breakdown of the headline batch call
mkdir -p gpurun_out
for tag in "" occ3; do
  lib=genomics_rs_b200/libgxalign${tag:+_$tag}.so
  echo "== ${tag:-default}"
  GX_LIB_PATH=$PWD/$lib timeout 300 python bench.py --workload nw1m --steps 3 --warmup 3 --no-k0 > gpurun_out/bench_nw1m_${tag:-def}.json 2>&1; python tools/show_bench.py gpurun_out/bench_nw1m_${tag:-def}.json | head -1
  GX_LIB_PATH=$PWD/$lib GX_K=4 timeout 300 python bench.py --workload nw1m --steps 3 --warmup 3 --no-k0 > gpurun_out/bench_nw1m_k4_${tag:-def}.json 2>&1; python tools/show_bench.py gpurun_out/bench_nw1m_k4_${tag:-def}.json | head -1
done
timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -12

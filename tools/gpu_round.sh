#!/bin/bash
# one gpurun call: GPU tests, then the bench lines of every workload (outputs under gpurun_out/)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for wl in corona45 brca2_global brca2_local; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 $( [ $wl != corona45 ] && echo --no-cpu-baseline --no-k0 ) > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
done
timeout 600 python bench.py --workload reads150 --pairs 10000000 --steps 5 --warmup 3 --no-cpu-baseline --no-k0 > gpurun_out/bench_reads10m.json 2> gpurun_out/bench_reads10m.err
for k in 8 16; do
  GX_K=$k timeout 600 python bench.py --workload nw1m --steps 3 --warmup 3 --no-k0 > gpurun_out/bench_nw1m_k$k.json 2> gpurun_out/bench_nw1m_k$k.err
done
tail -c 600 gpurun_out/bench_*.json

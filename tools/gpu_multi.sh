#!/bin/bash
# multi-GPU gpurun call: banded multi-GPU test, then bench at N = $1 GPUs
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_banded.py -m gpu -x -q -k multi 2>&1 | tail -30 | tee gpurun_out/pytest_multi_$N.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
run --workload nw1m --steps 3 --warmup 2 --no-k0 > gpurun_out/bench_nw1m_n$N.json 2> gpurun_out/bench_nw1m_n$N.err
run --workload corona45 --steps 10 --warmup 3 --no-k0 > gpurun_out/bench_corona45_n$N.json 2> gpurun_out/bench_corona45_n$N.err
run --workload reads150 --pairs 10000000 --steps 5 --warmup 3 --no-k0 > gpurun_out/bench_reads10m_n$N.json 2> gpurun_out/bench_reads10m_n$N.err
tail -c 1500 gpurun_out/bench_*_n$N.json; tail -n 5 gpurun_out/bench_nw1m_n$N.err

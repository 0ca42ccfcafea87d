#!/bin/bash
# multi-GPU gpurun call: banded multi-GPU test, then the all-configs bench line at N = $1 GPUs (one process per GPU)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_banded.py -m gpu -x -q -k multi 2>&1 | tail -5 | tee gpurun_out/pytest_multi_$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/bench_all_n$N.json 2> gpurun_out/bench_all_n$N.err
tail -c 800 gpurun_out/bench_all_n$N.err
python tools/show_bench.py gpurun_out/bench_all_n$N.json

#!/bin/bash
for k in 4 8; do
GX_K=$k timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload nw1m --steps 3 --warmup 3 --no-k0 > gpurun_out/bench_nw1m_n2_k$k.json 2>/dev/null
python tools/show_bench.py gpurun_out/bench_nw1m_n2_k$k.json | head -1
GX_K=$k timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload corona45 --steps 5 --warmup 3 --no-k0 > gpurun_out/bench_c45_n2_k$k.json 2>/dev/null
python tools/show_bench.py gpurun_out/bench_c45_n2_k$k.json | head -1
done

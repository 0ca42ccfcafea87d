#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err
tail -c 600 gpurun_out/bench_all_n1.err; python tools/show_bench.py gpurun_out/bench_all_n1.json
GX_WALK_STATS=1 timeout 300 python tools/walk_stats.py 2>&1 | tail -12
bash tools/gpu_band_ab.sh 2>&1 | tail -26

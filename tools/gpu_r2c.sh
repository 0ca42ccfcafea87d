#!/bin/bash
# round 2, call C: GPU tests, all-configs bench, strip slack experiment, one full corona pair on the host
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err
tail -c 1500 gpurun_out/bench_all_n1.err; python tools/show_bench.py gpurun_out/bench_all_n1.json
for lead in 64 256 1024; do
  GX_START_LEAD=$lead timeout 300 python bench.py --workload corona45 --steps 5 --warmup 3 --no-cpu-baseline --no-k0 > gpurun_out/bench_corona45_lead$lead.json 2> gpurun_out/bench_corona45_lead$lead.err
  python tools/show_bench.py gpurun_out/bench_corona45_lead$lead.json
done
GX_WPC=1 timeout 300 python bench.py --workload corona45 --steps 5 --warmup 3 --no-cpu-baseline --no-k0 > gpurun_out/bench_corona45_wpc1.json 2>&1; python tools/show_bench.py gpurun_out/bench_corona45_wpc1.json
(timeout 600 python tools/full_pair_oracle.py | tail -c 1200) 2>&1

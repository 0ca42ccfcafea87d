#!/bin/bash
# round 2, call A: issue-rate probe (IDP.4A), GPU tests with the (K, R) register tiles, timing sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
make -C tools -s k0_probe && timeout 120 ./tools/k0_probe > gpurun_out/r2_k0_probe.txt 2>&1
tail -25 gpurun_out/r2_k0_probe.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep_kr.jsonl
timeout 900 python tools/sweep_kr.py --workloads brca2_global,brca2_local,corona6,corona45,nw200k --steps 5 > gpurun_out/sweep_kr.log 2>&1
tail -5 gpurun_out/sweep_kr.log

#!/bin/bash
# ncu evidence for the default bench command (corona45) + a score-only capture (nw1m, shortened).
# The .ncu-rep files are converted to csv on the box (gpurun_out/ is capped at 64 MiB).
set -x
mkdir -p gpurun_out
conv() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-k0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_corona45.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gx_fill_kernel -s 3 -c 1 -f -o gpurun_out/prof_fill_corona45 $CMD > gpurun_out/ncu_full.log 2>&1
conv prof_fill_corona45
CMD2="python bench.py --workload nw1m --length 200000 --steps 1 --warmup 1 --no-k0"
GX_K=16 $CMD2 > gpurun_out/plain3.log 2>&1 &&
GX_K=16 ncu --set full --clock-control none --import-source on -k regex:gx_fill_kernel -s 1 -c 1 -f -o gpurun_out/prof_fill_nw200k $CMD2 > gpurun_out/ncu_full2.log 2>&1
conv prof_fill_nw200k
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gx_walk_kernel -s 3 -c 1 -f -o gpurun_out/prof_walk_corona45 $CMD > gpurun_out/ncu_full3.log 2>&1
conv prof_walk_corona45
ls -la gpurun_out

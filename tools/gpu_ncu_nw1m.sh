#!/bin/bash
mkdir -p gpurun_out
CMD2="python bench.py --workload nw1m --steps 1 --warmup 3 --no-k0"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gx_fill_kernel -s 6 -c 1 -f -o gpurun_out/r2_prof_fill_nw1m $CMD2 > gpurun_out/ncu_full2.log 2>&1
ncu -i gpurun_out/r2_prof_fill_nw1m.ncu-rep --page raw --csv > gpurun_out/r2_prof_fill_nw1m.raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_fill_nw1m.ncu-rep --page source --csv > gpurun_out/r2_prof_fill_nw1m.source.csv 2>/dev/null
rm -f gpurun_out/r2_prof_fill_nw1m.ncu-rep
tail -2 gpurun_out/ncu_full2.log

#!/bin/bash
# end-of-round ncu evidence for the default bench command: launch list + one full capture of the fill kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-k0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gx_fill_kernel -s 3 -c 1 -f -o gpurun_out/prof_fill_final $CMD > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_fill_final.ncu-rep --page raw --csv > gpurun_out/prof_fill_final.raw.csv 2>/dev/null
rm -f gpurun_out/prof_fill_final.ncu-rep
ls -la gpurun_out/launches_final.csv gpurun_out/prof_fill_final.raw.csv

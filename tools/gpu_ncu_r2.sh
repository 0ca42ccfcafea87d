#!/bin/bash
# ncu evidence for the default bench's dominant kernel (corona45 fill) and the score-only fill (200 kbp pair):
# launch list of the bench command + one --set full capture each, converted to csv on the box
mkdir -p gpurun_out
CMD="python bench.py --workload corona45 --steps 2 --warmup 3 --no-cpu-baseline --no-k0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_corona45.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gx_fill_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_fill_corona45 $CMD > gpurun_out/ncu_full.log 2>&1
CMD2="python bench.py --workload nw1m --length 200000 --steps 2 --warmup 3 --no-k0"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gx_fill_kernel -s 8 -c 1 -f -o gpurun_out/r2_prof_fill_nw200k $CMD2 > gpurun_out/ncu_full2.log 2>&1
for f in r2_prof_fill_corona45 r2_prof_fill_nw200k; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/$f.source.csv 2>/dev/null
done
ls -la gpurun_out/r2_* | head; tail -3 gpurun_out/ncu_full.log gpurun_out/ncu_full2.log

#!/bin/bash
mkdir -p gpurun_out
conv() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
for k in 4 16; do
  export GX_K=$k GX_CHAIN1=1
  python tools/one_strip.py 100000 1 > gpurun_out/one_strip_k$k.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gx_fill_kernel -s 1 -c 1 -f -o gpurun_out/prof_strip1_k$k python tools/one_strip.py 100000 1 > gpurun_out/ncu_strip1_k$k.log 2>&1
  conv prof_strip1_k$k
done
cat gpurun_out/one_strip_k*.log

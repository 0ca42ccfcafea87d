#!/bin/bash
mkdir -p gpurun_out
for tag in r1 cur; do
  lib=genomics_rs_b200/libgxalign_r1.so; [ $tag = cur ] && lib=genomics_rs_b200/libgxalign.so
  GX_LIB_PATH=$PWD/$lib python tools/local_probe.py > gpurun_out/plain_local_$tag.log 2>&1 &&
  GX_LIB_PATH=$PWD/$lib ncu --set full --clock-control none --import-source on -k regex:gx_fill_kernel -s 3 -c 1 -f -o gpurun_out/prof_local_$tag python tools/local_probe.py > gpurun_out/ncu_local_$tag.log 2>&1
  ncu -i gpurun_out/prof_local_$tag.ncu-rep --page raw --csv > gpurun_out/prof_local_$tag.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_local_$tag.ncu-rep --page source --csv > gpurun_out/prof_local_$tag.source.csv 2>/dev/null
  rm -f gpurun_out/prof_local_$tag.ncu-rep
  cat gpurun_out/plain_local_$tag.log
done

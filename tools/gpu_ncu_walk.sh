mkdir -p gpurun_out
CMD="python bench.py --workload corona45 --steps 2 --warmup 3 --no-cpu-baseline --no-k0"
$CMD > gpurun_out/plain_w.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gx_walk_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_walk_corona45 $CMD > gpurun_out/ncu_walk.log 2>&1
ncu -i gpurun_out/r2_prof_walk_corona45.ncu-rep --page raw --csv > gpurun_out/r2_prof_walk_corona45.raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_walk_corona45.ncu-rep --page source --csv > gpurun_out/r2_prof_walk_corona45.source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r2_prof_walk_corona45.*

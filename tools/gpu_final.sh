#!/bin/bash
# final evidence of round 2 on one GPU: tests, the default bench line, ncu launch list + full capture of the dominant kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err
python tools/show_bench.py gpurun_out/bench_all_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 700 gpurun_out/bench_reference.json
CMD="python bench.py --workload corona45 --steps 2 --warmup 3 --no-cpu-baseline --no-k0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_corona45.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gx_fill_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_fill_corona45 $CMD > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r2_prof_fill_corona45.ncu-rep --page raw --csv > gpurun_out/r2_prof_fill_corona45.raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_fill_corona45.ncu-rep --page source --csv > gpurun_out/r2_prof_fill_corona45.source.csv 2>/dev/null
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gx_walk_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_walk_corona45 $CMD > gpurun_out/ncu_full3.log 2>&1
ncu -i gpurun_out/r2_prof_walk_corona45.ncu-rep --page raw --csv > gpurun_out/r2_prof_walk_corona45.raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r2_prof_* gpurun_out/r2_launches_corona45.csv

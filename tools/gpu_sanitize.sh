#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md): bash tools/gpu_sanitize.sh memcheck|racecheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 600 python tools/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 2400 compute-sanitizer --tool $TOOL --print-limit 50 python tools/sanitize_small.py > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
echo "exit $?" >> gpurun_out/r2_sanitizer_$TOOL.log
tail -15 gpurun_out/r2_sanitizer_$TOOL.log

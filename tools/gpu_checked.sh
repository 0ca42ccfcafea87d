#!/bin/bash
# The GPU test-suite and the small-shape sweep against the CHECKED build of the library (device-side bounds assertions;
# compute-sanitizer is closed on this pool).  Build first: GX_BUILD_TAG=chk GX_BUILD_DEFS=-DGX_CHECKED python -m genomics_rs_b200.build
mkdir -p gpurun_out
export GX_LIB_PATH=$PWD/genomics_rs_b200/libgxalign_chk.so
python - <<'PY' | tee gpurun_out/r2_checked_build.log
from genomics_rs_b200 import _lib
print(_lib.load().gx_version().decode())
PY
timeout 900 python tools/sanitize_small.py 2>&1 | tail -3 | tee -a gpurun_out/r2_checked_build.log
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee -a gpurun_out/r2_checked_build.log

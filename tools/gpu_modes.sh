#!/bin/bash
for tag in r1 "" sel td tds; do
  [ -f genomics_rs_b200/libgxalign${tag:+_$tag}.so ] || continue
  echo "== ${tag:-current}"
  GX_LIB_PATH=$PWD/genomics_rs_b200/libgxalign${tag:+_$tag}.so python tools/mode_probe.py brca2 | grep "local=1"
done

#!/usr/bin/env python
"""Freezes the config-4 parity scores computed by the CPU oracle:
  parity  : every local SW score of the parity set's first 100 000 pairs (s2 = s1 with 1/16 substitutions)
  strided : every 100th pair of the 10 M-pair throughput set (independent uniform reads)
into tests/golden/config4_scores.npz (int16).  bench.py and the GPU tests compare against this file; the GPU tests
also recompute a slice with the oracle itself.

    python tests/golden/make_config4_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from genomics_rs_b200 import workloads as wl  # noqa: E402
from oracle import gxo  # noqa: E402

out = {}
blob, off1, len1, off2, len2 = wl.reads150(0, wl.CONFIG4_PARITY_PAIRS, parity_set=True)
out["parity"] = gxo.score_batch(blob, off1, len1, off2, len2, wl.CONFIG_TOML, True, n_threads=os.cpu_count() or 1)
idx = np.arange(0, 10_000_000, wl.CONFIG4_STRIDE, dtype=np.uint64)
blob, off1, len1, off2, len2 = wl.reads150_pairs(idx)
out["strided"] = gxo.score_batch(blob, off1, len1, off2, len2, wl.CONFIG_TOML, True, n_threads=os.cpu_count() or 1)
for k, v in out.items():
    assert v.min() >= 0 and v.max() < 32768
    print(k, v.size, "pairs; min/mean/max", int(v.min()), float(v.mean()), int(v.max()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "config4_scores.npz"), parity=out["parity"].astype(np.int16),
                    strided=out["strided"].astype(np.int16), stride=np.int64(wl.CONFIG4_STRIDE))

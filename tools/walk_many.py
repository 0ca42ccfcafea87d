#!/usr/bin/env python
"""Fill / walk time of traceback plans with MANY pairs (the walk runs one CTA per pair: residency matters, not the chain).
    [GX_LIB_PATH=genomics_rs_b200/libgxalign_other.so] python tools/walk_many.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genomics_rs_b200 as gx  # noqa: E402
from genomics_rs_b200 import _lib, workloads as wl  # noqa: E402

_lib.ensure_init(0)
rng = np.random.default_rng(7)


def mutated(n_pairs, length):
    pairs = []
    for _ in range(n_pairs):
        a = rng.integers(0, 4, size=length, dtype=np.uint8)
        b = a.copy()
        flip = rng.random(length) < 0.1
        b[flip] = rng.integers(0, 4, size=int(flip.sum()), dtype=np.uint8)
        cut = rng.integers(0, length, size=max(1, length // 50))
        b = np.delete(b, cut)
        pairs.append((np.frombuffer(b"ACGT", dtype=np.uint8)[a].tobytes(), np.frombuffer(b"ACGT", dtype=np.uint8)[b].tobytes()))
    return pairs


for n_pairs, length, local in [(8192, 150, False), (8192, 150, True), (2048, 1000, False), (512, 4000, False), (296, 8000, False), (300, 8000, False)]:
    blob, off1, len1, off2, len2 = gx.pack_pairs(mutated(n_pairs, length))
    plan = gx.Plan(len1, len2, wl.CONFIG_TOML, local, traceback=True)
    plan.upload(blob, off1, off2)
    for _ in range(3):
        plan.execute()
    f, w = [], []
    for _ in range(7):
        plan.execute()
        f.append(plan.fill_ms)
        w.append(plan.walk_ms)
    res, ops, off = plan.fetch()
    print(f"{n_pairs:6d} x {length:5d} local={int(local)}: fill {np.median(f):8.3f} ms  walk {np.median(w):8.3f} ms  K {int(plan.stat(15))}  ops {int(res['n_ops'].sum())} score-sum {int(res['score'].sum())}", flush=True)
    plan.close()

#!/usr/bin/env python
"""Small-shape run of every kernel for compute-sanitizer (memcheck / racecheck on the inter-strip hand-off,
the band links, the walk windows and the read kernels).  Every result is checked against the oracle, so a sanitizer
run is also a parity run.

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import genomics_rs_b200 as gx  # noqa: E402
from genomics_rs_b200 import _lib  # noqa: E402
from oracle import gxo  # noqa: E402
from conftest import random_pair  # noqa: E402

SC = (1, -2, -1, -5)
_lib.ensure_init(0)
rng = np.random.default_rng(7)
dims = [(0, 0), (1, 1), (5, 130), (33, 257), (300, 520), (700, 129), (4100, 140), (150, 1100)]
pairs = [random_pair(rng, m, n, similar=bool(k % 2)) for k, (m, n) in enumerate(dims)]
n_checked = 0
for k in (4, 8, 16):
    for chain1 in (0, 1):
        os.environ["GX_K"], os.environ["GX_CHAIN1"] = str(k), str(chain1)
        for is_local in (False, True):
            got = gx.align_batch(pairs, SC, is_local)
            for (a, b), r in zip(pairs, got):
                o = gxo.align_linear(a, b, SC, is_local)
                assert r.score == o.score and np.array_equal(r.ops, o.ops) and tuple(r.start) == tuple(o.start), (k, chain1, is_local, len(a), len(b))
                n_checked += 1
        # ticket mode (more strips than the forced grid), score only + start cell
        os.environ["GX_TICKETS"] = "1"
        got = gx.align_batch(pairs, SC, True, traceback=False, start_cell=True)
        for (a, b), r in zip(pairs, got):
            sc, si, sj = gxo.score_linear(a, b, SC, True)
            assert r.score == sc and tuple(r.start) == (si, sj)
        os.environ.pop("GX_TICKETS")
        # 3 column bands emulated in one kernel (the multi-GPU decomposition), incl. a band edge inside a strip
        a, b = random_pair(rng, 900, 1500)
        assert gx.nw_score_banded_local(a, b, SC, 3) == gxo.score_linear(a, b, SC, False)[0]
for var in ("GX_K", "GX_CHAIN1"):
    os.environ.pop(var, None)
# LCS-at-first-max pass, planes kernel, read kernels (32-bit and s16x2)
got = gx.align_batch(pairs[:6], SC, True, lcs_at_max=True)
for (a, b), r in zip(pairs[:6], got):
    assert r.matches_at_max == gxo.align_faithful(a, b, SC, True).lcs_at_first_max
for g_, e_ in zip(gx.score_planes(pairs[3][0][:30], pairs[3][1][:60], SC, False), gxo.planes(pairs[3][0][:30], pairs[3][1][:60], SC, False)):
    assert np.array_equal(g_, e_)
reads = [random_pair(rng, int(rng.integers(0, 152)), int(rng.integers(0, 152))) for _ in range(1500)]
blob, off1, len1, off2, len2 = gx.pack_pairs(reads)
for is_local in (True, False):
    assert np.array_equal(gx.score_batch(blob, off1, len1, off2, len2, SC, is_local), gxo.score_batch(blob, off1, len1, off2, len2, SC, is_local, n_threads=4))
print(f"sanitize_small ok: {n_checked} alignments + bands + LCS + planes + reads match the oracle")

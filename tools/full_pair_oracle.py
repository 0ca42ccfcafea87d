#!/usr/bin/env python
"""One FULL coronavirus pair through the oracle's FAITHFUL variant (the reference's algorithm as written: 48-byte cells,
column-major (m+1) x (n+1) table = ~43 GB, zero-filled first, i-outer / j-inner, stateless retrace) on the host cores of
a GPU box.  Two purposes (missing #4): (1) a like-for-like CPU baseline for config 3 -- the bench's bounded
samples use prefixes; (2) it pins the corona goldens, which were frozen from the linear-memory variant, to the faithful one.

    python tools/full_pair_oracle.py [a b]     # default pair (4, 7): Covid_Wuhan x MERS_2014_USA, the most gapped alignment
Writes gpurun_out/r2_cpu_full_pair.json (copied to profiles/ by hand).  Needs ~45 GB of free host memory; refuses otherwise.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genomics_rs_b200 import workloads as wl  # noqa: E402
from oracle import gxo  # noqa: E402


def mem_available_gb():
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable:"):
            return int(line.split()[1]) / 1e6
    return 0.0


def main():
    a, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (4, 7)
    seqs, jobs = wl.corona_pairs()
    s1, s2 = seqs[a], seqs[b]
    cells = (len(s1) + 1) * (len(s2) + 1)
    need = cells * 48 / 1e9
    out = {"pair": [a, b], "names": [wl.CORONA[a], wl.CORONA[b]], "m": len(s1), "n": len(s2), "cells": cells, "table_gb": need,
           "mem_available_gb": mem_available_gb(), "nproc": os.cpu_count()}
    try:
        out["cpu_model"] = next(l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name"))
    except Exception:
        pass
    path = os.path.join(ROOT, "gpurun_out", "r2_cpu_full_pair.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if out["mem_available_gb"] < need + 6:
        out["skipped"] = "not enough free host memory for the 48 B/cell table"
        json.dump(out, open(path, "w"), indent=1)
        print(json.dumps(out))
        return
    so = gxo.build(march="native")
    t0 = time.perf_counter()
    r = gxo.align_faithful(s1, s2, wl.CONFIG_TOML, False, so=so)
    dt = time.perf_counter() - t0
    gold = next(c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_goldens.json")))["corona"] if c["pair"] == [a, b])
    got = {"score": r.score, "start": list(r.start), "end": list(r.end), "n_ops": int(len(r.ops)), "matches": r.matches,
           "mismatches": r.mismatches, "gap_extensions": r.gap_extensions, "opening_gaps": r.opening_gaps,
           "op_hash": "%016x" % gxo.hash_ops(r.ops, r.start), "first_max": list(r.first_max), "lcs_at_first_max": r.lcs_at_first_max}
    out.update({"seconds": dt, "fill_s": r.fill_ms / 1e3, "walk_s": r.walk_ms / 1e3, "gcups": cells / dt / 1e9, "cores": 1,
                "what": "oracle faithful variant = the reference's alignment_table + retrace as written, one thread, -O3 -march=native",
                "result": got, "equals_golden": all(got[k] == gold[k] for k in got if k in gold),
                "golden_variant_was": gold.get("variant")})
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))
    if not out["equals_golden"]:
        raise SystemExit("faithful variant differs from the frozen golden")


if __name__ == "__main__":
    main()

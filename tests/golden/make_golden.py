#!/usr/bin/env python
"""Regenerates tests/golden/* .  Run in the build container only (needs /root/reference).

Outputs
  fasta/*.fasta.gz          the reference's public FASTA *inputs* for the alignment path (test_data/ and
                            comparison_data/), gzip-ed unchanged; the GPU box has no /root/reference.
  reference_vectors.json    the three golden vectors of the reference's own tests/test_alignment.rs
                            (transcribed by hand: inputs, TEST_CONFIG, expected counters and op list).
  oracle_goldens.json       outputs of oracle/gx_oracle.c on every fixture: faithful variant up to BRCA2,
                            linear variant for the 45 coronavirus pairs (the faithful table would be 43 GB).
                            These are ORACLE outputs, not reference outputs (the reference cannot run here);
                            they freeze the oracle so that later edits cannot silently change behaviour.
"""
import gzip
import json
import os
import sys
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import gxo  # noqa: E402

CONFIG_TOML = (1, -2, -1, -5)   # /root/reference/config.toml:1-5  (s_match, s_mismatch, g, h)
TEST_CONFIG = (1, -2, -2, -5)   # /root/reference/tests/test_alignment.rs:4-11

PAIR_FIXTURES = ["test1", "test2_short", "test3_short", "test4", "Opsin1_colorblindness_gene", "Human-Mouse-BRCA2-cds"]
CORONA = sorted(f[:-6] for f in os.listdir(os.path.join(REF, "comparison_data")) if f.endswith(".fasta"))


def read_fasta(path):
    """sequence.rs:45-95 restated (host-side helper for the generator only)."""
    seqs, have = [], False
    with open(path) as fh:
        for line in fh.read().split("\n"):
            line = line.rstrip("\r")
            if not line:
                continue
            if line.startswith(">"):
                seqs.append([line[1:].strip(), ""])
                have = True
            elif have:
                seqs[-1][1] += line.strip()
    return seqs


def summarize(r):
    return dict(score=r.score, start=list(r.start), end=list(r.end), n_ops=int(len(r.ops)), matches=r.matches,
                mismatches=r.mismatches, gap_extensions=r.gap_extensions, opening_gaps=r.opening_gaps,
                first_max=list(r.first_max), lcs_at_first_max=r.lcs_at_first_max,
                op_hash="%016x" % gxo.hash_ops(r.ops, r.start))


def corona_job(args):
    a, b = args
    s1 = read_fasta(os.path.join(REF, "comparison_data", CORONA[a] + ".fasta"))[0][1]
    s2 = read_fasta(os.path.join(REF, "comparison_data", CORONA[b] + ".fasta"))[0][1]
    r = gxo.align_linear(s1, s2, CONFIG_TOML, False)
    d = summarize(r)
    d.update(pair=[a, b], m=len(s1), n=len(s2), variant="linear")
    return d


def main():
    os.makedirs(os.path.join(HERE, "fasta"), exist_ok=True)
    for name in PAIR_FIXTURES:
        raw = open(os.path.join(REF, "test_data", name + ".fasta"), "rb").read()
        with gzip.GzipFile(os.path.join(HERE, "fasta", name + ".fasta.gz"), "wb", mtime=0) as fh:
            fh.write(raw)
    for name in CORONA:
        raw = open(os.path.join(REF, "comparison_data", name + ".fasta"), "rb").read()
        with gzip.GzipFile(os.path.join(HERE, "fasta", name + ".fasta.gz"), "wb", mtime=0) as fh:
            fh.write(raw)

    ref_vectors = dict(
        source="/root/reference/tests/test_alignment.rs",
        scores=dict(s_match=1, s_mismatch=-2, g=-2, h=-5),
        is_local=False,
        cases=[
            dict(name="test_simple_matches", lines="23-53", s1="ACGT", s2="ACGT", score=4, matches=4, mismatches=0,
                 opening_gaps=0, gap_extensions=0,
                 alignment=[["Match", 4, 4], ["Match", 3, 3], ["Match", 2, 2], ["Match", 1, 1]]),
            dict(name="test_gaps", lines="55-90", s1="ACGT", s2="AGCGT", score=None, matches=3, mismatches=1,
                 opening_gaps=1, gap_extensions=0,
                 alignment=[["Match", 4, 5], ["Match", 3, 4], ["Match", 2, 3], ["OpenInsert", 1, 2], ["Mismatch", 1, 1]]),
            dict(name="test_affine_gap", lines="92-139", s1="ACGGATAAAAAAAATC", s2="ACGGATAAAATC", score=None, matches=12,
                 mismatches=0, opening_gaps=1, gap_extensions=3,
                 alignment=[["Match", 16, 12], ["Match", 15, 11], ["Match", 14, 10], ["Match", 13, 9], ["Match", 12, 8],
                            ["Match", 11, 7], ["OpenDelete", 10, 6], ["Delete", 9, 6], ["Delete", 8, 6], ["Delete", 7, 6],
                            ["Match", 6, 6], ["Match", 5, 5], ["Match", 4, 4], ["Match", 3, 3], ["Match", 2, 2], ["Match", 1, 1]]),
        ])
    json.dump(ref_vectors, open(os.path.join(HERE, "reference_vectors.json"), "w"), indent=1)

    goldens = dict(scores=dict(s_match=1, s_mismatch=-2, g=-1, h=-5), note="oracle outputs, config.toml scoring",
                   pairs=[], corona_order=CORONA, corona=[])
    for name in PAIR_FIXTURES:
        s = read_fasta(os.path.join(REF, "test_data", name + ".fasta"))
        for is_local in (False, True):
            r = gxo.align_faithful(s[0][1], s[1][1], CONFIG_TOML, is_local)
            l = gxo.align_linear(s[0][1], s[1][1], CONFIG_TOML, is_local)
            assert summarize(r) == summarize(l), (name, is_local)
            d = summarize(r)
            d.update(fixture=name, is_local=is_local, m=len(s[0][1]), n=len(s[1][1]), variant="faithful==linear",
                     faithful_fill_ms=round(r.fill_ms, 1), faithful_walk_ms=round(r.walk_ms, 1))
            goldens["pairs"].append(d)
            print(d, flush=True)
    jobs = [(a, b) for a in range(len(CORONA)) for b in range(a + 1, len(CORONA))]
    with Pool(8) as pool:
        for d in pool.imap(corona_job, jobs):
            goldens["corona"].append(d)
            print(d["pair"], d["score"], d["n_ops"], d["op_hash"], flush=True)
    json.dump(goldens, open(os.path.join(HERE, "oracle_goldens.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

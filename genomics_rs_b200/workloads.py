"""The five BASELINE.json workloads as (pairs / blobs) generators.  Host-side only; SURVEY.md 8d fixes the
synthetic streams (splitmix64) so that every rank and the oracle see identical bytes."""
from __future__ import annotations

import gzip
import os
from typing import List, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FASTA_DIR = os.path.join(ROOT, "tests", "golden", "fasta")
CONFIG_TOML = (1, -2, -1, -5)   # the reference's config.toml:1-5

CORONA = ["Covid_Australia", "Covid_Brazil", "Covid_India", "Covid_USA-CA4", "Covid_Wuhan", "MERS_2012_KF600620",
          "MERS_2014_KY581694", "MERS_2014_USA_KP223131", "SARS_2003_GU553363", "SARS_2017_MK062179"]

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_LUT = np.frombuffer(b"ACGT", np.uint8)


def read_fasta_gz(name: str) -> List[Tuple[str, str]]:
    from .sequence import SequenceContainer
    import tempfile
    raw = gzip.open(os.path.join(FASTA_DIR, name + ".fasta.gz"), "rb").read()
    with tempfile.NamedTemporaryFile(suffix=".fasta", delete=False) as fh:
        fh.write(raw)
        path = fh.name
    try:
        sc = SequenceContainer()
        sc.from_fasta(path)
    finally:
        os.unlink(path)
    return [(s.name, s.sequence) for s in sc.sequences]


def splitmix64(seed: np.ndarray, k: np.ndarray) -> np.ndarray:
    """k-th output (k >= 1) of splitmix64 streams started at `seed` (vectorised, wrapping uint64)."""
    with np.errstate(over="ignore"):
        z = (seed + k * _GAMMA) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def bases_from_streams(seeds: np.ndarray, length: int) -> np.ndarray:
    """[len(seeds), length] base codes 0..3: each 64-bit output yields 32 bases, 2 bits LSB-first."""
    nw = (length + 31) // 32
    k = np.arange(1, nw + 1, dtype=np.uint64)[None, :]
    words = splitmix64(seeds.astype(np.uint64)[:, None], k)                  # [n, nw]
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, None, :]
    codes = ((words[:, :, None] >> shifts) & np.uint64(3)).astype(np.uint8)  # [n, nw, 32]
    return codes.reshape(len(seeds), nw * 32)[:, :length]


def corona_pairs():
    """config 3: all-vs-all of the 10 genomes in sorted-filename order, (a,b) with a<b, s1=a."""
    seqs = [read_fasta_gz(n)[0][1].encode() for n in CORONA]
    jobs = [(a, b) for a in range(len(seqs)) for b in range(a + 1, len(seqs))]
    return seqs, jobs


def brca2_pair():
    s = read_fasta_gz("Human-Mouse-BRCA2-cds")
    return s[0][1].encode(), s[1][1].encode()


def _reads_block(p: np.ndarray, length: int, parity_set: bool):
    """bases (as ACGT bytes) of the pairs with indices p: ([len(p), length], [len(p), length])"""
    seeds1 = np.uint64(0x5EED0150) + np.uint64(2) * p.astype(np.uint64)
    c1 = bases_from_streams(seeds1, length)
    if parity_set:
        # the q=1 stream supplies one random byte per base: low nibble == 0 (probability 1/16) substitutes the base
        nw = (length * 8 + 63) // 64
        k = np.arange(1, nw + 1, dtype=np.uint64)[None, :]
        w = splitmix64((seeds1 + np.uint64(1))[:, None], k)
        by = ((w[:, :, None] >> (np.arange(8, dtype=np.uint64) * np.uint64(8))[None, None, :]) & np.uint64(0xFF)).astype(np.uint8)
        by = by.reshape(len(p), nw * 8)[:, :length]
        sub = (by & 0x0F) == 0
        r = (by >> 4) % 3
        c2 = np.where(sub, (c1 + 1 + r) % 4, c1).astype(np.uint8)
    else:
        c2 = bases_from_streams(seeds1 + np.uint64(1), length)
    return _LUT[c1], _LUT[c2]


def reads150_pairs(indices, length: int = 150, parity_set: bool = False, chunk: int = 1 << 17, threads: int = 0):
    """config 4: pair p = (stream 0x5EED0150+2p, stream 0x5EED0150+2p+1), `length` bases each, for an arbitrary
    array of pair indices.  parity_set: s2 := s1 with each base substituted with probability 1/16 (from the q=1 stream).
    Returns (blob uint8 [n*2*length], off1, len1, off2, len2)."""
    indices = np.ascontiguousarray(indices, np.uint64)
    n_pairs = int(indices.size)
    blob = np.empty((n_pairs, 2, length), np.uint8)

    def fill(s):
        e = min(n_pairs, s + chunk)
        a, b = _reads_block(indices[s:e], length, parity_set)
        blob[s:e, 0, :] = a
        blob[s:e, 1, :] = b

    starts = list(range(0, n_pairs, chunk))
    threads = threads or min(8, os.cpu_count() or 1)
    if threads > 1 and len(starts) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:     # numpy releases the GIL inside the big element-wise kernels
            list(ex.map(fill, starts))
    else:
        for s in starts:
            fill(s)
    base = np.arange(n_pairs, dtype=np.uint64) * np.uint64(2 * length)
    lens = np.full(n_pairs, length, np.uint64)
    return blob.reshape(-1), base, lens, base + np.uint64(length), lens.copy()


def reads150(first_pair: int, n_pairs: int, length: int = 150, parity_set: bool = False, chunk: int = 1 << 17):
    """config 4: the contiguous range of pairs [first_pair, first_pair + n_pairs) (see reads150_pairs)."""
    return reads150_pairs(np.arange(first_pair, first_pair + n_pairs, dtype=np.uint64), length, parity_set, chunk)


# SURVEY 8d, config 4 parity: every score of the parity set's first 100 000 pairs and a strided 1 % of the throughput set
CONFIG4_PARITY_PAIRS = 100_000
CONFIG4_STRIDE = 100


def long_pair(length: int = 1_000_000):
    """config 5: s1 uniform from stream 0x5EED1000; s2 = s1 with per-base substitution probability 1/32."""
    c1 = bases_from_streams(np.array([0x5EED1000], np.uint64), length)[0]
    nw = (length + 7) // 8
    w = splitmix64(np.array([0x5EED1001], np.uint64)[:, None], np.arange(1, nw + 1, dtype=np.uint64)[None, :])
    by = ((w[:, :, None] >> (np.arange(8, dtype=np.uint64) * np.uint64(8))[None, None, :]) & np.uint64(0xFF)).astype(np.uint8)
    by = by.reshape(-1)[:length]
    sub = (by & 0x1F) == 0
    r = (by >> 5) % 3
    c2 = np.where(sub, (c1 + 1 + r) % 4, c1).astype(np.uint8)
    return _LUT[c1], _LUT[c2]


def lpt_shards(costs: List[int], n_shards: int) -> List[List[int]]:
    """Longest-processing-time greedy assignment of items (by cost) to n_shards bins."""
    order = sorted(range(len(costs)), key=lambda k: (-costs[k], k))
    bins: List[List[int]] = [[] for _ in range(n_shards)]
    load = [0] * n_shards
    for k in order:
        b = min(range(n_shards), key=lambda x: (load[x], x))
        bins[b].append(k)
        load[b] += costs[k]
    return [sorted(b) for b in bins]

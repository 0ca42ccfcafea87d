import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

CONFIG_TOML = (1, -2, -1, -5)   # reference config.toml
TEST_CONFIG = (1, -2, -2, -5)   # reference tests/test_alignment.rs:4-11
# (K, R) register tiles the fill kernel is built with (GX_COMBOS in csrc/gx_api.cu): K columns x R rows per lane per step
KR_COMBOS = [(4, 1), (8, 1), (16, 1)]


def force_kr(monkeypatch, k, r):
    monkeypatch.setenv("GX_K", str(k))
    monkeypatch.setenv("GX_R", str(r))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def read_fasta_gz(name):
    """[(name, sequence)] of tests/golden/fasta/<name>.fasta.gz, parsed like sequence.rs:45-95."""
    raw = gzip.open(os.path.join(GOLDEN, "fasta", name + ".fasta.gz"), "rb").read().decode()
    seqs = []
    for line in raw.split("\n"):
        line = line.rstrip("\r")
        if not line:
            continue
        if line.startswith(">"):
            seqs.append([line[1:].strip(), ""])
        elif seqs:
            seqs[-1][1] += line.strip()
    return [(a, b) for a, b in seqs]


@pytest.fixture(scope="session")
def goldens():
    return json.load(open(os.path.join(GOLDEN, "oracle_goldens.json")))


@pytest.fixture(scope="session")
def ref_vectors():
    return json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import gxo
    gxo.lib()
    return gxo


def random_pair(rng, m, n, alphabet=b"ACGT", similar=True, sub=0.15, indel=0.05):
    a = rng.choice(np.frombuffer(alphabet, np.uint8), size=m)
    if not similar:
        return a, rng.choice(np.frombuffer(alphabet, np.uint8), size=n)
    out = []
    i = 0
    while len(out) < n:
        r = rng.random()
        if i >= m or r < indel:
            out.append(rng.choice(np.frombuffer(alphabet, np.uint8)))
        elif r < 2 * indel:
            i += 1
        else:
            c = a[i]
            if rng.random() < sub:
                c = rng.choice(np.frombuffer(alphabet, np.uint8))
            out.append(c)
            i += 1
    return a, np.array(out[:n], np.uint8)

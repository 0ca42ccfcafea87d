"""Column-banded global NW (BASELINE config 5): host logic on CPU (band ranges, the SPMD driver over gloo with an
oracle-backed band), the band decomposition of the oracle itself against the frozen config-5 prefix scores, and --
marked gpu -- the CUDA band path against the oracle and those goldens."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import CONFIG_TOML, GOLDEN, KR_COMBOS, TEST_CONFIG, force_kr, random_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# ------------------------------------------------------------------------------------------------ CPU
def test_band_range_partitions_every_table():
    from genomics_rs_b200 import band_range
    for n in (1, 7, 8, 9, 100, 4095, 4096 * 8, 4096 * 8 + 1, 65536, 262144, 999_999, 1_000_000, 8191 * 16, 8192 * 16 + 5):
        for bands in (1, 2, 3, 4, 8, 16):
            if n < bands:
                continue
            r = [band_range(n, bands, b) for b in range(bands)]
            assert r[0][0] == 0 and r[-1][0] + r[-1][1] == n
            assert all(w > 0 for _, w in r), (n, bands, r)
            assert all(r[k][0] + r[k][1] == r[k + 1][0] for k in range(bands - 1))
            if n // bands >= max(4096, 512 * bands):      # wide tables: whole strips
                assert all(c0 % 512 == 0 for c0, _ in r)


def test_oracle_bands_reproduce_config5_prefix(oracle):
    """the band decomposition (left boundary = (V, I) of the previous band's last column) is exact:
    4096 x 4096 prefix of config 5, frozen score, through 1, 3 and 8 oracle bands"""
    from genomics_rs_b200 import workloads as wl, band_range
    gold = json.load(open(os.path.join(GOLDEN, "config5_scores.json")))
    a, b = wl.long_pair(4096)
    for bands in (1, 3, 8):
        left, sc = None, None
        for k in range(bands):
            c0, w = band_range(4096, bands, k)
            sc, left = oracle.nw_band(a, b[c0:c0 + w], c0, CONFIG_TOML, left)
        assert sc == gold["prefix_scores"]["4096"]
    assert oracle.score_linear(a, b, CONFIG_TOML, False)[0] == gold["prefix_scores"]["4096"]


class _OracleBand:
    """CPU stand-in for genomics_rs_b200.Band used by the gloo test: same interface, the boundary column travels
    by dist.send/recv between the ranks the exchanged handles name."""

    def __init__(self, m, n_total, n_bands, first, last, scores):
        from genomics_rs_b200 import band_range
        assert last == first + 1
        self.m, self.band, self.n_bands, self.scores = m, first, n_bands, scores
        self.col0, self.width = band_range(n_total, n_bands, first)
        self.left = self.right = None
        self.executes = 0

    def export(self):
        return b"rank%03d" % dist.get_rank() + bytes(57)

    def connect(self, left, right):
        self.left = int(left[4:7]) if left is not None else None
        self.right = int(right[4:7]) if right is not None else None

    def upload(self, s1, s2):
        self.s1, self.s2 = np.asarray(s1), np.asarray(s2)[self.col0:self.col0 + self.width]

    def execute(self):
        import torch
        from oracle import gxo
        left = None
        if self.left is not None:
            t = torch.zeros(2, self.m, dtype=torch.int64)
            dist.recv(t, src=self.left)
            left = (t[0].numpy(), t[1].numpy())
        self._score, (v, i) = gxo.nw_band(self.s1, self.s2, self.col0, self.scores, left)
        if self.right is not None:
            dist.send(torch.from_numpy(np.stack([v, i])), dst=self.right)
        self.executes += 1

    def score(self):
        return self._score if self.band == self.n_bands - 1 else None

    def close(self):
        pass


def _spmd_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genomics_rs_b200 import nw_score_banded
    from oracle import gxo
    rng = np.random.default_rng(3)
    a, b = random_pair(rng, 700, 900)
    sc, band = nw_score_banded(a, b, CONFIG_TOML, band_factory=_OracleBand, steps=2)
    ret[rank] = (sc, band.left, band.right, band.executes, gxo.score_linear(a, b, CONFIG_TOML, False)[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_spmd_driver_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_spmd_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        sc, left, right, n_exec, exp = ret[r]
        assert sc == exp                                  # every rank learns the score of the last band
        assert left == (r - 1 if r > 0 else None) and right == (r + 1 if r + 1 < world else None)
        assert n_exec == 2


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gx():
    import genomics_rs_b200 as gx
    from genomics_rs_b200 import _lib
    _lib.ensure_init()
    return gx


@pytest.mark.gpu
@pytest.mark.parametrize("k,r", KR_COMBOS)
def test_banded_local_vs_oracle(gx, oracle, k, r, monkeypatch):
    """N bands emulated on one GPU (one kernel over all bands' strips) against the oracle: sizes that put band
    edges inside strips, on panel boundaries and next to the table edge"""
    force_kr(monkeypatch, k, r)
    rng = np.random.default_rng(21)
    cases = [(1, 1, 1), (5, 9, 3), (300, 8, 8), (700, 900, 2), (4096, 1000, 3), (4097, 1025, 5), (9000, 5000, 8),
             (100, 40000, 4), (12000, 33000, 8), (20000, 4096 * 8, 8), (8200, 70000, 16)]
    for scores in (CONFIG_TOML, TEST_CONFIG):
        for m, n, bands in cases:
            a, b = random_pair(rng, m, n, similar=bool((m + n) % 2))
            exp = oracle.score_linear(a, b, scores, False)[0]
            assert gx.nw_score_banded_local(a, b, scores, bands) == exp, (m, n, bands, scores)
    assert gx.nw_score_banded_local(b"", b"ACGT", CONFIG_TOML, 2) == -5 - 4
    assert gx.nw_score_banded_local(b"ACG", b"", CONFIG_TOML, 1) == -5 - 3
    assert gx.nw_score_banded_local(b"", b"", CONFIG_TOML, 1) == 0


@pytest.mark.gpu
def test_band_reexecute(gx, oracle):
    rng = np.random.default_rng(4)
    a, b = random_pair(rng, 9000, 14000)
    exp = oracle.score_linear(a, b, CONFIG_TOML, False)[0]
    band = gx.Band(len(a), len(b), 4, 0, 4, CONFIG_TOML)
    band.upload(a, b)
    for _ in range(5):
        band.execute()
        assert band.score() == exp
    assert band.stat(16) == 5
    band.close()


@pytest.mark.gpu
@pytest.mark.parametrize("prefix,bands", [(4096, 8), (65536, 1), (65536, 8), (262144, 4), (1_000_000, 1), (1_000_000, 8)])
def test_config5_prefixes(gx, prefix, bands):
    """BASELINE config 5 (splitmix64 pair) against the scores frozen from the CPU oracle
    (tools/freeze_config5.py), up to the full 1 Mbp x 1 Mbp table (1e12 cells)"""
    from genomics_rs_b200 import workloads as wl
    gold = json.load(open(os.path.join(GOLDEN, "config5_scores.json")))
    a, b = wl.long_pair(1_000_000)
    assert gx.nw_score_banded_local(a[:prefix], b[:prefix], CONFIG_TOML, bands) == gold["prefix_scores"][str(prefix)]


@pytest.mark.gpu
def test_banded_multi_gpu():
    """one process per GPU, boundary columns stored into the neighbour's HBM over NVLink (needs >= 2 GPUs)"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "banded_multi_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "banded multi-gpu ok" in out.stdout

"""world_size-2 gloo test (CPU) of the multi-GPU host logic: LPT sharding covers every pair exactly once,
rank 0 gets all results in input order.  The aligner is injected (CPU oracle) so no GPU is needed."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from genomics_rs_b200.scatter import scatter_align, shard_indices  # noqa: E402
from genomics_rs_b200.workloads import lpt_shards  # noqa: E402


def test_lpt_shards_partition():
    rng = np.random.default_rng(0)
    costs = [int(x) for x in rng.integers(1, 1000, size=45)]
    for world in (1, 2, 4, 8):
        bins = lpt_shards(costs, world)
        assert sorted(k for b in bins for k in b) == list(range(45))
        loads = [sum(costs[k] for k in b) for b in bins]
        assert max(loads) - min(loads) <= max(costs)
    assert shard_indices([(3, 4)] * 5, 2) == [[0, 2, 4], [1, 3]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gxo
    rng = np.random.default_rng(5)
    lut = np.frombuffer(b"ACGT", np.uint8)
    pairs = [(lut[rng.integers(0, 4, size=int(rng.integers(1, 300)))].tobytes(),
              lut[rng.integers(0, 4, size=int(rng.integers(1, 300)))].tobytes()) for _ in range(17)]

    def cpu_align(p, scores, is_local, traceback):
        return [gxo.align_linear(a, b, scores, is_local).score for a, b in p]

    out = scatter_align(pairs, (1, -2, -1, -5), False, align_fn=cpu_align)
    if rank == 0:
        exp = [gxo.align_linear(a, b, (1, -2, -1, -5), False).score for a, b in pairs]
        ret["ok"] = (out == exp)
    else:
        ret[f"none{rank}"] = out is None
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_align_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok") is True and ret.get("none1") is True

"""torchrun worker of tests/test_banded.py::test_banded_multi_gpu (one process per GPU)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import genomics_rs_b200 as gx
    from genomics_rs_b200 import _lib, workloads as wl
    from conftest import CONFIG_TOML, TEST_CONFIG, random_pair
    from oracle import gxo
    _lib.ensure_init(local)
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(8)
    for m, n, scores in [(3000, 5000, CONFIG_TOML), (9000, 40000, TEST_CONFIG), (20000, 70000, CONFIG_TOML)]:
        a, b = random_pair(rng, m, n)
        sc, band = gx.nw_score_banded(a, b, scores, steps=3)      # 3 executes: parity + ack flow control
        exp = gxo.score_linear(a, b, scores, False)[0]
        assert sc == exp, (rank, m, n, sc, exp)
        band.close()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "config5_scores.json")))
    a, b = wl.long_pair(1_000_000)
    for prefix in (65536, 262144, 1_000_000):
        sc, band = gx.nw_score_banded(a[:prefix], b[:prefix], CONFIG_TOML, steps=2)
        assert sc == gold["prefix_scores"][str(prefix)], (rank, prefix, sc)
        if rank == 0:
            print(f"prefix {prefix}: score {sc}, rank-0 band fill {band.fill_ms:.2f} ms", flush=True)
        band.close()
    dist.barrier()
    if rank == 0:
        print(f"banded multi-gpu ok (world {world})", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

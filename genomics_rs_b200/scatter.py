"""Many-pair workloads over several GPUs: one process per GPU, pairs dealt by cost, no data-path collective
(SURVEY.md 8e).  torch.distributed is plumbing only: it carries the small per-pair results back."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

from .workloads import lpt_shards


def shard_indices(lengths: Sequence[Tuple[int, int]], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of pairs to ranks by table size (m+1)(n+1)."""
    costs = [(int(m) + 1) * (int(n) + 1) for m, n in lengths]
    return lpt_shards(costs, world)


def scatter_align(pairs: Sequence[Tuple[object, object]], scores, is_local: bool, traceback: bool = True,
                  align_fn: Optional[Callable] = None, group=None) -> Optional[list]:
    """SPMD: every rank calls this with the same `pairs`; each aligns its own shard on its own GPU
    (gx_align_batch) and rank 0 receives all results in input order (other ranks get None).
    `align_fn(pairs, scores, is_local, traceback) -> list` can be injected (tests use the CPU oracle)."""
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    shards = shard_indices([(len(a), len(b)) for a, b in pairs], world)
    mine = shards[rank]
    if align_fn is None:
        from .alignment import align_batch
        align_fn = lambda p, s, l, t: align_batch(p, s, l, traceback=t)  # noqa: E731
    local = align_fn([pairs[k] for k in mine], scores, is_local, traceback) if mine else []
    if world == 1:
        out = [None] * len(pairs)
        for k, r in zip(mine, local):
            out[k] = r
        return out
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(list(zip(mine, local)), gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * len(pairs)
    for part in gathered:
        for k, r in part:
            out[k] = r
    return out

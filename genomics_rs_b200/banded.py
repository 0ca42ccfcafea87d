"""One very long pair, global score only, cut into column bands over the GPUs of a node (BASELINE config 5,
SURVEY.md 8e).  Replaces `alignment_table` (algo.rs:151-282) for tables the reference could never allocate.

One process per GPU.  Rank r owns band r of `world` bands; the fill kernel of band r stores the (E,I) of its
last column straight into GPU r+1's HBM (CUDA IPC mapping of the neighbour's link block, NVLink) and GPU r+1's
kernel polls its own memory -- the boundary exchange is fused into the fill kernel, there is no collective call
on the data path.  torch.distributed is plumbing only: it carries the 64-byte IPC handles once and the final score.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .alignment import _as_u8, _scores_struct


def band_range(n_total: int, n_bands: int, band: int) -> Tuple[int, int]:
    """(col0, width) of band `band`: gx_band_range (pure host arithmetic, identical on every rank)."""
    lib = _lib.load()
    c0, w = C.c_uint64(), C.c_uint64()
    _lib.check(lib.gx_band_range(int(n_total), int(n_bands), int(band), C.byref(c0), C.byref(w)))
    return int(c0.value), int(w.value)


class Band:
    """gx_band wrapper: this process's bands [first, last) of an (m+1) x (n+1) global table."""

    HANDLE_BYTES = 64

    def __init__(self, m: int, n_total: int, n_bands: int, first: int, last: int, scores, device: Optional[int] = None):
        self.lib = _lib.ensure_init(device)
        self.m, self.n_total, self.n_bands, self.first, self.last = int(m), int(n_total), int(n_bands), int(first), int(last)
        self._h = C.c_void_p()
        _lib.check(self.lib.gx_band_create(self.m, self.n_total, self.n_bands, self.first, self.last,
                                           _scores_struct(scores), C.byref(self._h)))

    def export(self) -> bytes:
        buf = (C.c_uint8 * self.HANDLE_BYTES)()
        _lib.check(self.lib.gx_band_export(self._h, buf, self.HANDLE_BYTES))
        return bytes(buf)

    def connect(self, left: Optional[bytes], right: Optional[bytes]) -> None:
        lb = (C.c_uint8 * self.HANDLE_BYTES).from_buffer_copy(left) if left is not None else None
        rb = (C.c_uint8 * self.HANDLE_BYTES).from_buffer_copy(right) if right is not None else None
        _lib.check(self.lib.gx_band_connect(self._h, lb, rb))

    def upload(self, s1, s2) -> None:
        a, b = _as_u8(s1), _as_u8(s2)
        if a.size != self.m or b.size != self.n_total:
            raise ValueError("sequence lengths differ from the band's table")
        _lib.check(self.lib.gx_band_upload(self._h, a.ctypes.data if a.size else None, b.ctypes.data if b.size else None))

    def execute(self) -> None:
        _lib.check(self.lib.gx_band_execute(self._h))

    def score(self) -> Optional[int]:
        sc, valid = C.c_int64(), C.c_int()
        _lib.check(self.lib.gx_band_score(self._h, C.byref(sc), C.byref(valid)))
        return int(sc.value) if valid.value else None

    def stat(self, what: int) -> float:
        return float(self.lib.gx_band_stat(self._h, what))

    @property
    def fill_ms(self) -> float:
        return self.stat(0)

    def close(self) -> None:
        if self._h:
            self.lib.gx_band_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nw_score_banded_local(s1, s2, scores, n_bands: int = 1) -> int:
    """All bands on this process's GPU (gx_nw_score_banded): one kernel, the N-rank decomposition on one device."""
    lib = _lib.ensure_init()
    a, b = _as_u8(s1), _as_u8(s2)
    sc = C.c_int64()
    _lib.check(lib.gx_nw_score_banded(a.ctypes.data if a.size else None, a.size, b.ctypes.data if b.size else None, b.size,
                                      _scores_struct(scores), int(n_bands), C.byref(sc)))
    return int(sc.value)


def connect_ring(band, group=None) -> None:
    """Exchange the link-block handles of all ranks and map the two neighbours' blocks (collective)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    handles = [None] * world
    dist.all_gather_object(handles, band.export(), group=group)
    band.connect(handles[rank - 1] if rank > 0 else None, handles[rank + 1] if rank + 1 < world else None)
    dist.barrier(group=group)   # every inbox is mapped (and initialised) before any kernel may write into it


_BAND_CACHE = {}


def clear_cache() -> None:
    """close the bands kept by nw_score_banded(..., cache=True) (collective when they span ranks: call it on every rank)"""
    for band in _BAND_CACHE.values():
        band.close()
    _BAND_CACHE.clear()


def nw_score_banded(s1, s2, scores, group=None, band_factory=None, steps: int = 1, cache: bool = False):
    """SPMD: every rank calls this with the same full s1, s2.  Returns (score, band) on every rank; the caller
    owns `band` (band.execute() may be repeated, the same number of times on every rank; band.close()).
    cache=True keeps the connected band (plan, device buffers, the neighbours' CUDA-IPC mappings) for the next call with
    the same shape, scores and group -- a stream of equally shaped pairs then pays for creation and the handle exchange
    once; the returned band belongs to the cache (do not close it; clear_cache() does).
    `band_factory(m, n, world, rank, rank+1, scores)` can be injected (the CPU test uses an oracle-backed band)."""
    import torch
    import torch.distributed as dist
    a, b = _as_u8(s1), _as_u8(s2)
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    if world > 1 and 0 < b.size < world:
        raise ValueError("fewer columns than bands")
    make = band_factory or Band
    key = (int(a.size), int(b.size), world, rank, tuple(int(x) for x in (scores.as_tuple() if hasattr(scores, "as_tuple") else scores)),
           id(group), make)
    band = _BAND_CACHE.get(key) if cache else None
    if band is None:
        band = make(a.size, b.size, world, rank, rank + 1, scores)
        if world > 1:
            connect_ring(band, group)
        if cache:
            _BAND_CACHE[key] = band
    band.upload(a, b)
    for _ in range(steps):
        band.execute()
    sc = band.score()
    if world > 1:
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor([sc if sc is not None else 0], dtype=torch.int64, device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, world - 1) if group is not None else world - 1, group=group)
        sc = int(t.item())
    return sc, band

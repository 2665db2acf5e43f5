"""Multi-GPU plumbing of the hashRF path (one process per GPU, torch.distributed).

The matrix shards with no data-path collective: every output entry depends only on
the two trees' rows of the incidence, so rank r computes the aligned slice prf_shard_range(T, r, world)
of the packed triangle from the WHOLE bipartition table.  The only exchange is before the build:
ranks hold disjoint tree shards (the host extracted their bipartitions) and all-gather the keys
(NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def tree_shard(T: int, rank: int, world: int) -> Tuple[int, int]:
    """Trees [lo, hi) owned by `rank`: contiguous, balanced to within one tree."""
    return T * rank // world, T * (rank + 1) // world


def shard_rows(tree_off: np.ndarray, T: int, world: int) -> List[int]:
    """Number of bipartition entries each rank contributes (tree_off is the global T+1 offset array)."""
    return [int(tree_off[tree_shard(T, r, world)[1]] - tree_off[tree_shard(T, r, world)[0]]) for r in range(world)]


class KeyGather:
    """All-gathers per-rank [rows_r, W] int64 key blocks into one [nnz, W] table (rows_r may differ:
    blocks are padded to the longest).  Buffers are allocated once and reused every step."""

    def __init__(self, rows: List[int], W: int, device: torch.device):
        self.rows, self.W, self.world = rows, W, len(rows)
        self.max_rows = max(rows) if rows else 0
        self.full = torch.empty((sum(rows), W), dtype=torch.int64, device=device)
        self.gathered = torch.empty((self.world, self.max_rows, W), dtype=torch.int64, device=device)
        self.pad = torch.zeros((self.max_rows, W), dtype=torch.int64, device=device)

    def __call__(self, mine: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return mine
        self.pad[: mine.shape[0]].copy_(mine)
        dist.all_gather_into_tensor(self.gathered.view(-1), self.pad.view(-1))
        pos = 0
        for r in range(self.world):
            self.full[pos:pos + self.rows[r]].copy_(self.gathered[r, : self.rows[r]])
            pos += self.rows[r]
        return self.full


# ---------------------------------------------------------------------------------------------------
# Hash-partitioned build: every rank deduplicates one hash class of the bipartitions (prf_shard.cuh).
# The rank-local steps below are shared by the torch.distributed driver (sharded_load) and by the
# single-process loop-back driver the GPU tests use (sharded_load_loopback).
# ---------------------------------------------------------------------------------------------------
import ctypes as C  # noqa: E402
import os  # noqa: E402
import sys  # noqa: E402
import time  # noqa: E402

from ._native import prf_part, prf_stats, rf_lib  # noqa: E402
from .rfdistance import MEM_DEVICE, RFContext  # noqa: E402


class _DevView:
    """Zero-copy torch view of raw device memory (through __cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (ptr, False), "version": 3}


def _view(ptr: int, nbytes: int, device) -> torch.Tensor:
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevView(ptr, nbytes), device=device)[:nbytes]


def _partition(ctx: RFContext, keys: torch.Tensor, off: torch.Tensor, world: int):
    """keys [nnz_local, W] int64, off [T_local + 1] int64 (device).  Returns (send_keys grouped by class,
    rows per class, counts [world, T_local] int32)."""
    T_local, W = off.numel() - 1, keys.shape[1]
    send = torch.empty_like(keys)
    counts = torch.zeros((world, max(T_local, 1)), dtype=torch.int32, device=keys.device)
    first = (C.c_uint64 * (world + 1))()
    ctx._ck(ctx._L.prf_hashrf_partition(ctx._h, keys.data_ptr(), off.data_ptr(), T_local, W, world, send.data_ptr(),
                                        counts.data_ptr(), first))
    rows = [int(first[c + 1] - first[c]) for c in range(world)]
    return send, rows, counts[:, :T_local]


def _local_load(ctx: RFContext, recv_keys: torch.Tensor, counts_all: torch.Tensor, T: int, W: int) -> prf_part:
    """counts_all [T] int32: entries of every tree in my class, tree order == order of recv_keys."""
    off = torch.zeros(T + 1, dtype=torch.int64, device=recv_keys.device)
    torch.cumsum(counts_all.to(torch.int64), 0, out=off[1:])
    ctx._L.prf_debug_short_max(ctx._h, 0)   # the exported parts carry every list as suffix ranges
    try:
        ctx.hashrf_load(recv_keys.data_ptr(), off.data_ptr(), T, W, MEM_DEVICE)
    finally:
        ctx._L.prf_debug_short_max(ctx._h, 8)
    part = prf_part()
    ctx._ck(ctx._L.prf_hashrf_export(ctx._h, C.byref(part)))
    return part


_PART_ARRAYS = (("nb", 2, lambda p: p.n_trees), ("roff", 4, lambda p: p.n_trees + 1), ("ranges", 8, lambda p: p.n_ranges),
                ("members", 4, lambda p: p.n_members))
_PART_SCALARS = ("n_trees", "n_ranges", "n_members", "n_lists", "nnz", "n_unique", "n_hits")


def _merge(ctx: RFContext, T: int, W: int, parts: List[prf_part]) -> dict:
    arr = (prf_part * len(parts))(*parts)
    st = prf_stats()
    ctx._ck(ctx._L.prf_hashrf_merge(ctx._h, T, W, len(parts), arr, C.byref(st)))
    ctx.T, ctx.P = T, T * (T - 1) // 2 if T else 0
    return st.as_dict()


def sharded_load(ctx: RFContext, keys: torch.Tensor, off: torch.Tensor, T: int, stream=None) -> dict:
    """torch.distributed driver.  `keys`/`off` describe THIS rank's tree shard tree_shard(T, rank, world)
    (device tensors, int64).  On return ctx holds the merged structure: hashrf_compute works as usual."""
    world, rank, dev, W = dist.get_world_size(), dist.get_rank(), keys.device, keys.shape[1]
    trace = _Trace(rank == 0 and bool(os.environ.get("PRF_TRACE")), dev)
    with torch.cuda.stream(stream) if stream is not None else _null():
        send, rows, counts = _partition(ctx, keys, off, world)
        trace("partition")
        # rows each rank will receive from me / I receive from each rank
        in_rows = torch.tensor(rows, dtype=torch.int64, device=dev)
        out_rows = torch.empty_like(in_rows)
        dist.all_to_all_single(out_rows, in_rows)
        out_rows_h = out_rows.tolist()
        recv = torch.empty((sum(out_rows_h), W), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=out_rows_h, input_split_sizes=rows)
        trace("all-to-all keys")
        # per-tree counts: rank s sends me its trees' counts in my class; shards differ by at most one tree
        t_sizes = [tree_shard(T, r, world)[1] - tree_shard(T, r, world)[0] for r in range(world)]
        t_max = max(t_sizes)
        cs = torch.zeros((world, t_max), dtype=torch.int32, device=dev)
        cs[:, : counts.shape[1]] = counts
        cr = torch.empty_like(cs)
        dist.all_to_all_single(cr.view(-1), cs.view(-1))
        counts_all = torch.cat([cr[s, : t_sizes[s]] for s in range(world)])
        trace("all-to-all counts")
        part = _local_load(ctx, recv, counts_all, T, W)
        trace("local build")
        # all-gather the partial structures: sizes first (tiny), then ONE packed buffer per rank
        mine = torch.tensor([getattr(part, k) for k in _PART_SCALARS], dtype=torch.int64, device=dev)
        allsc = torch.empty((world, len(_PART_SCALARS)), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allsc.view(-1), mine)
        allsc_h = allsc.tolist()
        parts = [prf_part() for _ in range(world)]
        for r in range(world):
            for k, v in zip(_PART_SCALARS, allsc_h[r]):
                setattr(parts[r], k, int(v))
        # layout of a packed buffer: every array at a 256-byte aligned offset sized for the longest rank
        offs, total = [], 0
        for name, esz, count in _PART_ARRAYS:
            nmax = max(int(count(parts[r])) for r in range(world)) * esz
            offs.append(total)
            total += (nmax + 255) // 256 * 256
        total = max(total, 256)
        pack = torch.empty(total, dtype=torch.uint8, device=dev)
        for (name, esz, count), o in zip(_PART_ARRAYS, offs):
            src = _view(getattr(part, name) or 0, int(count(part)) * esz, dev)
            pack[o:o + src.numel()].copy_(src)
        gathered = torch.empty((world, total), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered.view(-1), pack)
        keep = [gathered]
        for r in range(world):
            for (name, esz, count), o in zip(_PART_ARRAYS, offs):
                setattr(parts[r], name, gathered[r, o:].data_ptr())
        trace("all-gather structures")
        st = _merge(ctx, T, W, parts)
        trace("merge")
        del keep
    return st


class _Trace:
    """PRF_TRACE=1: wall-clock marks (with a device sync) of the sharded build's phases on rank 0."""

    def __init__(self, on, dev):
        self.on, self.dev = on, dev
        if on:
            torch.cuda.synchronize(dev)
            self.t = time.perf_counter()

    def __call__(self, what):
        if self.on:
            torch.cuda.synchronize(self.dev)
            now = time.perf_counter()
            sys.stderr.write(f"[sharded_load] {what:24s} +{1e3 * (now - self.t):8.3f} ms\n")
            self.t = now


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def sharded_load_loopback(ctxs: List[RFContext], bips: np.ndarray, tree_off: np.ndarray, device="cuda:0") -> List[dict]:
    """Single-process emulation of sharded_load over len(ctxs) "ranks" on one GPU (tests): the same
    rank-local steps, with the collectives replaced by slicing.  Every ctx ends up with the merged
    structure."""
    world, T, W = len(ctxs), len(tree_off) - 1, bips.shape[1]
    dev = torch.device(device)
    sends, rows, counts = [], [], []
    for r, ctx in enumerate(ctxs):
        lo, hi = tree_shard(T, r, world)
        b0, b1 = int(tree_off[lo]), int(tree_off[hi])
        keys = torch.from_numpy(np.ascontiguousarray(bips[b0:b1]).view(np.int64)).to(dev).reshape(-1, W)
        off = torch.from_numpy((tree_off[lo:hi + 1] - tree_off[lo]).astype(np.int64)).to(dev)
        s, rw, c = _partition(ctx, keys, off, world)
        torch.cuda.synchronize()
        sends.append(s), rows.append(rw), counts.append(c)
    parts, keep = [], []
    for r, ctx in enumerate(ctxs):
        chunks = []
        for s in range(world):
            start = sum(rows[s][:r])
            chunks.append(sends[s][start:start + rows[s][r]])
        recv = torch.cat(chunks) if chunks else torch.empty((0, W), dtype=torch.int64, device=dev)
        counts_all = torch.cat([counts[s][r] for s in range(world)])
        part = _local_load(ctx, recv, counts_all, T, W)
        # detach from the context: the merge below replaces the context's own structure
        cp = prf_part()
        for k in _PART_SCALARS:
            setattr(cp, k, getattr(part, k))
        for name, esz, count in _PART_ARRAYS:
            t = _view(getattr(part, name) or 0, int(count(part)) * esz, dev).clone()
            keep.append(t)
            setattr(cp, name, t.data_ptr() if t.numel() else 0)
        torch.cuda.synchronize()
        parts.append(cp)
        keep.append(recv)
    out = [_merge(ctx, T, W, parts) for ctx in ctxs]
    torch.cuda.synchronize()
    del keep
    return out

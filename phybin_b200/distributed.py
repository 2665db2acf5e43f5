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
# Hash-partitioned build over the ranks' GPUs (prf_team.cuh): torch.distributed only carries the 64-byte
# arena handles once; the data path is the library's own kernels storing into the peers' memory over NVLink.
# ---------------------------------------------------------------------------------------------------

def team_setup(ctx, T: int, W: int, nnz_total: int, max_local_entries: int, max_tree_entries: int) -> None:
    """Collective, once per problem shape: every rank allocates its exchange arena (prf_team_init) and opens the
    peers' (prf_team_connect).  `ctx` is an RFContext on this rank's GPU."""
    world, rank = dist.get_world_size(), dist.get_rank()
    handle = ctx.team_init(rank, world, T, W, nnz_total, max_local_entries, max_tree_entries)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    ctx.team_connect(b"".join(handles))
    dist.barrier()  # every rank has opened every arena before anybody writes into one


def team_load(ctx, keys: torch.Tensor, off: torch.Tensor, T: int) -> dict:
    """Collective: `keys` [nnz_local, W] int64 and `off` [T_local + 1] int64 describe THIS rank's trees
    tree_shard(T, rank, world) (device tensors visible to ctx's stream).  On return ctx holds the whole structure:
    ctx.hashrf_compute works on any range."""
    # the tensors may have been produced on torch's current stream: the library's stream waits for it first
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=keys.device)
    lib_stream.wait_stream(torch.cuda.current_stream(keys.device))
    return ctx.team_load(keys.data_ptr(), off.data_ptr(), off.numel() - 1, T)


# ---------------------------------------------------------------------------------------------------
# Flat clusters from a SHARDED threshold mask: every rank runs the union-find over its own shard,
# the label arrays are all-gathered (T x 4 bytes per rank) and merged by one more union-find on the device --
# the all-reduce(min) of labels without its iteration to a fixed point.
# ---------------------------------------------------------------------------------------------------

def gather_labels(labels: np.ndarray, device: torch.device | None = None) -> np.ndarray:
    """All-gathers every rank's label array: [world, T] uint32 (NCCL on `device`, gloo when None)."""
    world = dist.get_world_size()
    mine = torch.from_numpy(labels.astype(np.int64))
    if device is not None:
        mine = mine.to(device)
    out = torch.empty((world, mine.numel()), dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(out.view(-1), mine)
    return out.cpu().numpy().astype(np.uint32)


def sharded_components(ctx, p_begin: int, p_end: int, device: torch.device | None = None, d_mask: int = 0):
    """Collective: components of { d <= editdist } when rank r computed (and holds the mask of) the packed range
    [p_begin, p_end) = prf_shard_range(T, r, world).  Returns (labels[T], n_components), identical on every rank."""
    mine, _ = ctx.mask_components_shard(p_begin, p_end, d_mask=d_mask)
    return ctx.mask_components_shard(0, 0, seeds=gather_labels(mine, device))

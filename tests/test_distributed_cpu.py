"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: tree sharding, the key all-gather and the
aligned split of the packed triangle.  No compute calls: the kernels need a GPU."""
import ctypes
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phybin_b200 import _native
from phybin_b200.distributed import KeyGather, gather_labels, shard_rows, tree_shard
from phybin_b200.host import Forest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        f = Forest.synthetic(T, n, seed=42, p_missing=0.1)  # ragged per-tree counts
        bips, off = f.all_bips()
        lo, hi = tree_shard(T, rank, world)
        mine = torch.from_numpy(np.ascontiguousarray(bips[int(off[lo]):int(off[hi])]).view(np.int64))
        rows = shard_rows(off, T, world)
        g = KeyGather(rows, bips.shape[1], torch.device("cpu"))
        for _ in range(2):  # buffers are reused across steps
            full = g(mine)
        ok = np.array_equal(full.numpy().view(np.uint64), bips)
        # every rank's slice of the packed triangle, from the C ABI helper
        L = _native.rf_lib()
        b, e = ctypes.c_uint64(), ctypes.c_uint64()
        assert L.prf_shard_range(T, rank, world, ctypes.byref(b), ctypes.byref(e)) == 0
        ranges = [None] * world
        dist.all_gather_object(ranges, (b.value, e.value))
        # the exchange step of the sharded flat clusters: every rank's label array, in rank order, on every rank
        labs = gather_labels(np.arange(T, dtype=np.uint32) // (rank + 2))
        ok = ok and labs.shape == (world, T) and all(np.array_equal(labs[r], np.arange(T) // (r + 2)) for r in range(world))
        with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as fh:
            fh.write(f"{int(ok)} {ranges}")
    finally:
        dist.destroy_process_group()


def test_two_rank_key_gather_and_shard_ranges(tmp_path):
    T, n, world = 501, 40, 2
    mp.spawn(_worker, args=(world, _free_port(), T, n, str(tmp_path)), nprocs=world, join=True)
    P = T * (T - 1) // 2
    for r in range(world):
        ok, ranges = open(tmp_path / f"r{r}.txt").read().split(" ", 1)
        assert ok == "1"
        ranges = eval(ranges)
        assert ranges[0][0] == 0 and ranges[-1][1] == P
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


def test_tree_shards_tile_the_forest():
    for T in (1, 7, 100000):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = tree_shard(T, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == T

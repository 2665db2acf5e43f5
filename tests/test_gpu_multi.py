"""Multi-GPU build (prf_team.cuh) exercised on ONE GPU: prf_multi with every rank on device 0 (single-process teams
synchronise with stream events, so several ranks can share a device), against the single-context result and the
literal oracle.  The flag-kernel barrier of the one-process-per-GPU mode needs distinct GPUs: bench.py --gpus N."""
import numpy as np
import pytest

from oracle import oracle as orc
from phybin_b200.host import Forest
from phybin_b200.rfdistance import MultiRF, RFContext
from tests.util import random_multifurcating_newicks

pytestmark = pytest.mark.gpu
NNI_FAMILY = 1


def _single(bips, off, editdist=-1):
    ctx = RFContext(0)
    try:
        return ctx.hashrf(bips, off, editdist)
    finally:
        ctx.close()


@pytest.mark.parametrize("world,T,n,family", [(1, 200, 20, 0), (2, 300, 30, 0), (3, 700, 64, 0), (4, 513, 40, NNI_FAMILY),
                                              (8, 1000, 100, 0), (5, 37, 12, 0)])
def test_multi_matches_single_context_and_oracle(world, T, n, family):
    f = Forest.synthetic(T, n, 0x5EED0100 + world, family, 6, 0.0)
    bips, off = f.all_bips()
    m = MultiRF(devices=[0] * world)
    try:
        out, _, st = m.hashrf(bips, off)
        out2, _, st2 = m.hashrf(bips, off)   # the arenas are reused
    finally:
        m.close()
    ref, _, st1 = _single(bips, off)
    assert np.array_equal(out, ref) and np.array_equal(out2, ref)
    assert st["n_unique"] == st1["n_unique"] and st["n_shared"] == st1["n_shared"]
    want = orc.lower_to_packed_upper(orc.Forest([f.to_newick()]).hashrf(n, "literal"), T)
    assert np.array_equal(out.astype(np.int64), want)


def test_multi_nonuniform_duplicates_and_mask():
    import random
    f = Forest.parse([random_multifurcating_newicks(random.Random(11), 400, 24)])
    bips, off = f.all_bips()
    # list some bipartitions twice inside a tree: a tree's entries are a SET
    rows, offs = [], [0]
    for t in range(len(off) - 1):
        b = bips[int(off[t]):int(off[t + 1])]
        parts = [b, b[:2]] if t % 3 == 0 else [b]
        rows += parts
        offs.append(offs[-1] + sum(len(p) for p in parts))
    bips2 = np.concatenate(rows)
    off2 = np.array(offs, dtype=np.uint64)
    ref, mref, _ = _single(bips, off, 6)
    m = MultiRF(devices=[0, 0, 0])
    try:
        out, mask, _ = m.hashrf(bips2, off2, 6)
    finally:
        m.close()
    assert np.array_equal(out, ref)
    assert np.array_equal(mask, mref)


def test_team_of_one_through_the_staged_entry_points():
    import torch
    f = Forest.synthetic(500, 50, 77, 0, 6, 0.0)
    bips, off = f.all_bips()
    ref, _, _ = _single(bips, off)
    ctx = RFContext(0)
    try:
        T, W = len(off) - 1, bips.shape[1]
        ctx.team_init(0, 1, T, W, int(off[-1]), int(off[-1]), int(np.diff(off.astype(np.int64)).max()))
        ext = torch.cuda.ExternalStream(ctx.stream)
        with torch.cuda.stream(ext):
            d_b = torch.from_numpy(bips.view(np.int64)).cuda()
            d_o = torch.from_numpy(off.view(np.int64)).cuda()
        ext.synchronize()
        ctx.team_load(d_b.data_ptr(), d_o.data_ptr(), T, T)
        P = T * (T - 1) // 2
        ctx.hashrf_compute(0, P)
        out = np.empty(P, dtype=np.uint16)
        ctx._ck(ctx._L.prf_fetch(ctx._h, 0, P, out.ctypes.data))
        assert np.array_equal(out, ref)
    finally:
        ctx.close()


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
def test_sharded_mask_components_equal_whole_mask_components(world):
    """the mask is sharded like the matrix; per-shard union-find + merge of the label arrays gives the
    components of the whole graph (prf_mask_components_shard, prf_multi_components)."""
    from phybin_b200.rfdistance import shard_range
    T, thr = 900, 6
    f = Forest.synthetic(T, 40, 0x5EED0200 + world, NNI_FAMILY, 6, 0.0)
    bips, off = f.all_bips()
    ctx = RFContext(0)
    try:
        ctx.hashrf_load(bips, off)
        ctx.hashrf_compute(editdist=thr)
        want, n_want = ctx.mask_components()
        assert 1 < n_want < T
        # shard by shard on one context: each range computed (context-owned mask of exactly that range), then merged
        parts = []
        for r in range(world):
            b, e = shard_range(T, r, world)
            ctx.hashrf_compute(b, e, editdist=thr)
            parts.append(ctx.mask_components_shard(b, e)[0])
        got, n_got = ctx.mask_components_shard(0, 0, seeds=np.stack(parts))
        assert np.array_equal(got, want) and n_got == n_want
        # chained form: a shard's components seeded with the previous shards' labels
        acc = None
        for r in range(world):
            b, e = shard_range(T, r, world)
            ctx.hashrf_compute(b, e, editdist=thr)
            acc, n_acc = ctx.mask_components_shard(b, e, seeds=None if acc is None else acc[None, :])
        assert np.array_equal(acc, want) and n_acc == n_want
    finally:
        ctx.close()
    m = MultiRF(devices=[0] * world)
    try:
        m.hashrf(bips, off, thr)
        got, n_got = m.components()
    finally:
        m.close()
    assert np.array_equal(got, want) and n_got == n_want

"""Randomised GPU stress of the hashRF path (run on a B200; not part of pytest because of its length):
random shapes / families / shard ranges / thresholds, every kernel variant and tile configuration against
the closed-form oracle, plus repeated runs of the same input to expose races (results must be identical).
usage: python scripts/stress_gpu.py [seconds]"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as orc
from phybin_b200.host import Forest, NNI_FAMILY
from phybin_b200.rfdistance import RFContext, shard_range
from tests.util import oracle_packed, random_multifurcating_newicks

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = random.Random(int(os.environ.get("STRESS_SEED", "12345")))
ctxs = {False: RFContext(0), True: RFContext(0)}
ctxs[True].debug_small_tiles(True)
t_end = time.time() + budget
n_cases = n_checks = 0
while time.time() < t_end:
    kind = rng.choice(["random", "nni", "multi"])
    T = rng.choice([2, 3, 17, 64, 257, 700, 1500, 2300])
    n = rng.choice([4, 9, 16, 33, 64, 65, 100, 130])
    seed = rng.randrange(1 << 30)
    if kind == "multi":
        text = [random_multifurcating_newicks(random.Random(seed), T, max(n, 5))]
        f = Forest.parse(text)
        o = orc.Forest(text)
    else:
        f = Forest.synthetic(T, n, seed=seed, family=NNI_FAMILY if kind == "nni" else 0, nni_max=rng.choice([1, 3, 8]))
        o = orc.Forest([f.to_newick()])
    want = oracle_packed(o, "hashrf", "closed")
    bips, off = f.all_bips()
    T = len(f)
    P = T * (T - 1) // 2
    thr = int(rng.choice([0, 2, int(np.median(want)) if P else 0, 10 ** 6]))
    for small in (False, True):
        ctx = ctxs[small]
        ctx.hashrf_load(bips, off)
        for variant in (1, 2):
            world = rng.choice([1, 1, 2, 3, 5])
            for rep in range(2 if small else 1):   # small tiles: many single-tile rows per CTA -> repeat to catch races
                for r in range(world):
                    b, e = shard_range(T, r, world)
                    ctx.hashrf_compute(b, e, editdist=thr, variant=variant)
                    got = ctx.fetch(b, e)
                    assert np.array_equal(got.astype(np.int64), want[b:e]), (kind, T, n, seed, small, variant, world, r)
                    m = np.unpackbits(ctx.fetch_mask(b, e).view(np.uint8), bitorder="little")[: e - b].astype(bool)
                    assert np.array_equal(m, want[b:e] <= thr), ("mask", kind, T, n, seed, small, variant, world, r)
                    n_checks += 1
        if P:
            ctx.hashrf_compute(0, P, editdist=thr)
            labels, ncomp = ctx.mask_components()
            assert ncomp == len(set(labels.tolist()))
    n_cases += 1
print(f"stress ok: {n_cases} inputs, {n_checks} (variant, tile, shard) checks in {budget:.0f} s")

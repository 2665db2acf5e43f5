"""Randomised GPU stress of the rows around the matrix (run on a B200; not part of pytest because of its length):
device extraction (allBips / raw clades) against the host extraction word for word and against the oracle's
allBips, one-shot from trees against the staged calls, strict consensus of random partitions and of mask components
against the oracle, flat clusters of every linkage against the naive restatement, the device agglomeration
(prf_agglomerate: merge lists, bit for bit) against the host restatement on tie-heavy random matrices, and the tolerant
matrix of forests with repeated leaf labels (prf_tolerant_load_dups) against the oracle's literal naiveDistMatrix.
usage: python scripts/stress_consumers.py [seconds]"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as orc
from oracle import pyref
from phybin_b200.host import Forest, NNI_FAMILY, bits_to_labels
from phybin_b200.rfdistance import EXTRACT_BIPS, EXTRACT_CLADES, RFContext, packed_index
from phybin_b200 import host as phost
from tests.test_tolerant_model_cpu import dup_newicks
from tests.util import oracle_packed, random_multifurcating_newicks

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
rng = random.Random(int(os.environ.get("STRESS_SEED", "777")))
ctx = RFContext(0)
t_end = time.time() + budget
n_cases = n_cons = n_clu = n_agg = n_dup = 0
while time.time() < t_end:
    kind = rng.choice(["random", "nni", "multi", "multi_missing"])
    T = rng.choice([1, 2, 3, 17, 64, 257, 600])
    n = rng.choice([3, 4, 5, 9, 16, 33, 64, 65, 100, 130, 260])
    seed = rng.randrange(1 << 30)
    if kind.startswith("multi"):
        text = [random_multifurcating_newicks(random.Random(seed), T, n, p_missing=0.25 if kind == "multi_missing" else 0.0,
                                              unary=rng.choice([0.0, 0.2]))]
        f = Forest.parse(text)
        o = orc.Forest(text)
    else:
        f = Forest.synthetic(T, max(n, 3), seed=seed, family=NNI_FAMILY if kind == "nni" else 0, nni_max=rng.choice([1, 3, 8]),
                             p_missing=rng.choice([0.0, 0.0, 0.1]))
        o = orc.Forest([f.to_newick()])
    T = len(f)
    W = f.words + rng.choice([0, 0, 1])
    # extraction: device == host, word for word, both modes; device == oracle as sets
    hb, ho = f.all_bips(W)
    ctx.allbips(f, EXTRACT_BIPS, W)
    db, do = ctx.allbips_fetch()
    assert np.array_equal(do, ho) and np.array_equal(db, hb), ("bips", kind, T, n, seed, W)
    for t in rng.sample(range(T), min(T, 5)):
        assert sorted(bits_to_labels(k) for k in db[int(do[t]):int(do[t + 1])]) == sorted(o.all_bips(t)), ("oracle bips", kind, T, n, seed, t)
    hc, hco, hl = f.raw_clades(W)
    ctx.allbips(f, EXTRACT_CLADES, W)
    dc, dco, dl = ctx.allbips_fetch()
    assert np.array_equal(dco, hco) and np.array_equal(dc, hc) and np.array_equal(dl, hl), ("clades", kind, T, n, seed, W)
    # one-shot from trees == staged calls on host arrays
    full = not f.has_duplicate_leaves()
    if T >= 2 and full:
        up, _, _ = ctx.hashrf_trees(f, tolerant=True)
        want, _, _ = ctx.tolerant(hc, hco, hl)
        assert np.array_equal(up, want), ("tolerant one-shot", kind, T, n, seed)
    if T >= 2:
        thr = rng.choice([0, 2, 4, 8])
        up, mask, _ = ctx.hashrf_trees(f, editdist=thr)
        want, wmask, _ = ctx.hashrf(hb, ho, editdist=thr)
        assert np.array_equal(up, want) and np.array_equal(mask, wmask), ("hashrf one-shot", kind, T, n, seed)
        # consensus: mask components and a random partition, against the oracle's intersection
        labels, _ = ctx.mask_components()
        group = np.array([rng.randrange(5) for _ in range(T)])
        labels2 = np.array([int(np.nonzero(group == group[t])[0][0]) for t in range(T)], dtype=np.uint32)
        for lab in (labels, labels2):
            keys, off, _ = ctx.consensus(lab)
            for c in np.unique(lab)[:6]:
                got = sorted(bits_to_labels(k) for k in keys[int(off[c]):int(off[c + 1])])
                assert got == o.consensus_bips(np.nonzero(lab == c)[0].tolist()), ("consensus", kind, T, n, seed, int(c))
                n_cons += 1
        if T <= 64:
            d = lambda i, j: int(up[packed_index(T, i, j)])
            for linkage in ("single", "complete", "UPGMA"):
                got, _ = ctx.flat_clusters(thr, linkage)
                assert got.tolist() == pyref.flat_clusters(d, T, linkage, thr), ("clusters", kind, T, n, seed, linkage, thr)
                n_clu += 1
                got_dev, _ = ctx.flat_clusters(thr, linkage, device_min=3)
                assert np.array_equal(got_dev, got), ("clusters on the device", kind, T, n, seed, linkage, thr)
    # the agglomeration itself: an arbitrary tie-heavy matrix on the device, merge list == host restatement
    import torch
    m = rng.choice([2, 3, 5, 31, 100, 333, 800])
    upper = np.random.default_rng(seed).integers(0, rng.choice([1, 2, 5, 40]) + 1, size=m * (m - 1) // 2).astype(np.uint16)
    t_up = torch.from_numpy(upper.astype(np.int16)).to("cuda:0")
    torch.cuda.synchronize()
    ctx.T = m
    for linkage in ("single", "complete", "UPGMA"):
        thresh = rng.choice([float("inf"), 0.0, 1.0, 2.5])
        hl, ha, hb_, hd = phost.agglomerate(upper, m, linkage, thresh)
        gl, ga, gb, gd, _ = ctx.agglomerate(None, linkage, thresh, t_up.data_ptr())
        assert np.array_equal(ga, ha) and np.array_equal(gb, hb_) and np.array_equal(gd, hd) and np.array_equal(gl, hl), ("agglo", m, seed, linkage, thresh)
        n_agg += 1
    del t_up
    # repeated leaf labels in --tolerant
    Td, nd = rng.choice([2, 9, 40]), rng.choice([2, 3, 4, 6, 20, 64, 65, 140])
    text = dup_newicks(random.Random(seed), Td, nd, p_missing=rng.choice([0.0, 0.2, 0.5]), p_dup=rng.choice([0.03, 0.3]))
    fd = Forest.parse([text])
    ctx.tolerant_load_dups(*fd.raw_clades_ex())
    ctx.tolerant_compute()
    assert np.array_equal(ctx.fetch().astype(np.int64), oracle_packed(orc.Forest([text]), "naive")), ("tolerant dups", Td, nd, seed)
    n_dup += 1
    n_cases += 1
print(f"agglomerations {n_agg}, repeated-label forests {n_dup};", end=" ")
print(f"consumer stress ok: {n_cases} inputs, {n_cons} consensus sets, {n_clu} clusterings in {budget:.0f} s")

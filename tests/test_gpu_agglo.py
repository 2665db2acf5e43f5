"""GPU tests of prf_agglomerate (C.dendrogram linkage trees dist, PhyBin.hs:358 + sliceDendro :415-422):
the device's merge list against the host restatement phb_agglomerate (same rule, same IEEE arithmetic => identical
merges and bitwise-equal distances), against scipy where there are no ties, and dendrogram.txt's text."""
import numpy as np
import pytest

from phybin_b200 import host
from phybin_b200.host import Forest, NNI_FAMILY
from phybin_b200.rfdistance import PhyBinRFError, RFContext, packed_index

pytestmark = pytest.mark.gpu
LINK = ("single", "complete", "UPGMA")


@pytest.fixture(scope="module")
def ctx():
    c = RFContext(0)
    yield c
    c.close()


def _load_matrix(ctx, upper: np.ndarray, T: int):
    """Puts an arbitrary packed uint16 triangle on the device; returns the torch tensor that owns it."""
    import torch
    t = torch.from_numpy(upper.astype(np.int16)).to("cuda:0")
    torch.cuda.synchronize()
    ctx.T = T
    return t


@pytest.mark.parametrize("m,hi", [(2, 3), (3, 1), (17, 2), (64, 3), (257, 4), (700, 2), (1500, 6), (3000, 9)])
def test_device_merges_equal_host_on_tie_heavy_matrices(ctx, m, hi):
    """Few distinct values => ties at every step; sizes cover 1 CTA .. 148 CTAs with several rows each."""
    rng = np.random.default_rng(m)
    upper = rng.integers(0, hi + 1, size=m * (m - 1) // 2).astype(np.uint16)
    t = _load_matrix(ctx, upper, m)
    for linkage in LINK:
        for thresh in (float("inf"), 1.0):
            hl, ha, hb, hd = host.agglomerate(upper, m, linkage, thresh)
            gl, ga, gb, gd, info = ctx.agglomerate(None, linkage, thresh, t.data_ptr())
            assert np.array_equal(ga, ha) and np.array_equal(gb, hb), (linkage, thresh)
            assert np.array_equal(gd, hd) and np.array_equal(gl, hl)
            assert info["n_merges"] == len(ha) and info["n_clusters"] == m - len(ha) and info["gpu_launches"] == 2


def test_device_agglomeration_of_a_subset_and_of_real_rf_distances(ctx):
    T, thr = 1200, 6
    f = Forest.synthetic(T, 40, 77, family=NNI_FAMILY, nni_max=6)
    ctx.hashrf_load_trees(f)
    ctx.hashrf_compute(editdist=thr)
    rng = np.random.default_rng(5)
    ids = np.sort(rng.choice(T, size=500, replace=False)).astype(np.uint32)
    sub = ctx.fetch_submatrix(ids)
    for linkage in LINK:
        for thresh in (float("inf"), float(thr), 0.0):
            hl, ha, hb, hd = host.agglomerate(sub, len(ids), linkage, thresh)
            gl, ga, gb, gd, _ = ctx.agglomerate(ids, linkage, thresh)
            assert np.array_equal(ga, ha) and np.array_equal(gb, hb) and np.array_equal(gd, hd) and np.array_equal(gl, hl)
    # flat clusters: every component on the device == every component on the host
    for linkage in ("complete", "UPGMA"):
        dev, kd = ctx.flat_clusters(thr, linkage, device_min=3)
        hst, kh = ctx.flat_clusters(thr, linkage, device_min=1 << 30)
        assert np.array_equal(dev, hst) and kd == kh
    # dendrogram.txt of every tree
    names = ["tree%d.tr" % i for i in range(T)]
    txt = ctx.dendrogram(names, "UPGMA")
    _, ha, hb, hd = host.agglomerate(ctx.fetch(), T, "UPGMA")
    assert txt == host.dendrogram_text(names, ha, hb, hd) and txt.count("Leaf") == T


def test_device_dendrogram_heights_equal_scipy_without_ties(ctx):
    """Distinct distances: the dendrogram is unique, so scipy's linkage heights are an independent check."""
    from scipy.cluster.hierarchy import linkage as sp_linkage
    m = 300
    P = m * (m - 1) // 2
    upper = (np.random.default_rng(3).permutation(P) + 1).astype(np.uint16)   # all distinct, < 65535
    t = _load_matrix(ctx, upper, m)
    for ours, theirs in (("single", "single"), ("complete", "complete"), ("UPGMA", "average")):
        _, ga, gb, gd, _ = ctx.agglomerate(None, ours, float("inf"), t.data_ptr())
        Z = sp_linkage(upper.astype(np.float64), theirs)
        assert len(gd) == m - 1 and np.allclose(gd, Z[:, 2], rtol=1e-12, atol=0)


def test_agglomerate_argument_errors(ctx):
    f = Forest.synthetic(50, 12, 3)
    ctx.hashrf_load_trees(f)
    ctx.hashrf_compute()
    with pytest.raises(PhyBinRFError):
        ctx.agglomerate(np.array([4, 2], dtype=np.uint32))          # not ascending
    with pytest.raises(PhyBinRFError):
        ctx.agglomerate(np.array([1, 50], dtype=np.uint32))         # id out of range
    with pytest.raises(PhyBinRFError):
        ctx._ck(ctx._L.prf_agglomerate(ctx._h, None, 50, None, 50, 3, 1.0, None, None, None, None, None))  # bad linkage
    labels, ma, mb, md, info = ctx.agglomerate(np.array([7], dtype=np.uint32))
    assert labels.tolist() == [0] and len(ma) == 0 and info["n_clusters"] == 1


# ---- --tolerant with repeated leaf labels (prf_tolerant_load_dups) ---------------------------------------------

@pytest.mark.parametrize("seed", range(8))
def test_tolerant_repeated_leaf_labels_vs_oracle(ctx, seed):
    """numLeaves counts a repeated label twice (CoreTypes.hs:289-291) while label sets stay sets: the DUPS
    instantiation of the tolerant kernel against the oracle's literal naiveDistMatrix."""
    import random
    from oracle import oracle as orc
    from tests.test_tolerant_model_cpu import dup_newicks
    from tests.util import oracle_packed
    rng = random.Random(500 + seed)
    n = rng.choice([2, 3, 4, 7, 30, 66, 130, 200])
    text = dup_newicks(rng, 40, n, p_missing=rng.choice([0.0, 0.15, 0.4]), p_dup=rng.choice([0.02, 0.3]))
    f, o = Forest.parse([text]), orc.Forest([text])
    assert f.has_duplicate_leaves()
    want = oracle_packed(o, "naive")
    ctx.tolerant_load_dups(*f.raw_clades_ex())
    ctx.tolerant_compute(editdist=3)
    assert np.array_equal(ctx.fetch().astype(np.int64), want)
    P = len(want)
    bits = np.unpackbits(ctx.fetch_mask().view(np.uint8), bitorder="little")[:P].astype(bool)
    assert np.array_equal(bits, want <= 3)
    # the public mirror routes there by itself; the one-shot device extraction refuses such trees loudly
    from phybin_b200.rfdistance import naiveDistMatrix
    mat, _ = naiveDistMatrix(f, ctx)
    assert np.array_equal(mat.upper.astype(np.int64), want)
    with pytest.raises(PhyBinRFError) as e:
        ctx.hashrf_trees(f, tolerant=True)
    assert e.value.code == -6


def test_tolerant_dups_form_equals_plain_form_without_repeats(ctx):
    import random
    from tests.util import random_multifurcating_newicks
    rng = random.Random(9)
    f = Forest.parse([random_multifurcating_newicks(rng, 300, 90, p_missing=0.2, unary=0.2)])
    want, _, _ = ctx.tolerant(*f.raw_clades())
    ctx.tolerant_load_dups(*f.raw_clades_ex())
    ctx.tolerant_compute()
    assert np.array_equal(ctx.fetch(), want)

"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the CPU oracle on the same inputs.  Bit-exact -- this is integer work."""
import ctypes as C
import io
import random

import numpy as np
import pytest

from oracle import oracle as orc
from phybin_b200 import _native
from phybin_b200.host import Forest, NNI_FAMILY, bits_to_labels
from phybin_b200.rfdistance import (EXTRACT_BIPS, EXTRACT_CLADES, DistanceMatrix, PhyBinRFError, RFContext, hashRF, naiveDistMatrix,
                                    packed_index, printDistMat, shard_range)
from tests.util import oracle_packed, random_multifurcating_newicks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["default", "small_blocks", "small_blocks_no_singles", "no_singles"])
def ctx(request):
    """Every test runs in four configurations of the build + sparse matrix kernel (prf_incidence.cuh, prf_rows.cuh):
      default                  production column blocks (8192 trees); lists of <= 8 trees expanded into singles
      small_blocks             256-tree debug blocks: small inputs exercise multi-block rows, partial first / last
                               tiles, and several passes over a row's lists (8 lists per pass)
      small_blocks_no_singles  same blocks, every shared bipartition is a bucketed list (no singles)
      no_singles               production blocks, no singles"""
    c = RFContext(0)
    c.debug_small_tiles(request.param.startswith("small_blocks"))
    if request.param.endswith("no_singles"):
        c._L.prf_debug_short_max(c._h, 1)
    yield c
    c.close()


# ---- primitives ---------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [1, 31, 4096, 4097, 100003, 1 << 21])
def test_scan_primitive(ctx, n):
    L = _native.rf_lib()
    rng = np.random.default_rng(n)
    x = rng.integers(0, 5, size=n, dtype=np.uint32)
    out = np.empty_like(x)
    tot = C.c_uint32(0)
    assert L.prf_debug_scan(ctx._h, x.ctypes.data, C.c_uint64(n), out.ctypes.data, C.byref(tot)) == 0
    want = np.concatenate([[0], np.cumsum(x, dtype=np.uint64)[:-1]]).astype(np.uint32)
    assert np.array_equal(out, want) and tot.value == int(x.sum())


@pytest.mark.parametrize("n,bits", [(2, 1), (1000, 8), (4096, 9), (50001, 17), (300000, 22)])
def test_sort_primitive_is_stable(ctx, n, bits):
    L = _native.rf_lib()
    rng = np.random.default_rng(n)
    k = rng.integers(0, 1 << bits, size=n, dtype=np.uint32)
    v = np.arange(n, dtype=np.uint32)
    k2, v2 = k.copy(), v.copy()
    assert L.prf_debug_sort(ctx._h, k2.ctypes.data, v2.ctypes.data, n, bits) == 0
    order = np.argsort(k, kind="stable")
    assert np.array_equal(k2, k[order]) and np.array_equal(v2, v[order])


# ---- golden fixtures ----------------------------------------------------------------------------

def test_reference_kat_t30_on_device(golden, ctx):
    c = golden["t30_literal"]
    mat, eachbips = naiveDistMatrix(Forest.parse(c["texts"], c["name_prefix"]), ctx)
    assert repr(mat) == "[[],[0],[1,0]]"  # TestAll.hs:31-32
    assert [sorted(list(b) for b in bs) for bs in eachbips] == [sorted(x) for x in c["all_bips"]]
    buf = io.StringIO()
    printDistMat(buf, mat)
    assert buf.getvalue().splitlines()[2:5] == ["0", "0 0", "1 0 0"]


@pytest.mark.parametrize("name", ["t30_file", "t1", "t2", "t4", "t5", "t3"])
def test_golden_hashrf_and_naive(golden, ctx, name):
    c = golden[name]
    f = Forest.parse(c["texts"], c["name_prefix"])
    mat, _ = naiveDistMatrix(f, ctx)
    assert mat.rows() == c["naive"]
    f.filter_num_leaves(len(c["labels"]))
    assert len(f) == c["hashrf_num_trees"]
    mat = hashRF(len(c["labels"]), f, ctx)
    assert mat.rows() == c["hashrf"]
    s = c.get("survey", {})
    if "sum" in s:
        assert int(mat.upper.astype(np.int64).sum()) == s["sum"] and int(mat.upper.max(initial=0)) == s["max"]


# ---- synthetic, against the literal oracle --------------------------------------------------------

def run_hashrf(ctx, f: Forest, editdist=-1):
    bips, off = f.all_bips()
    return ctx.hashrf(bips, off, editdist)


@pytest.mark.parametrize("T,n,family", [(100, 10, 0), (300, 30, 0), (700, 64, 0), (700, 65, 0),
                                        (400, 100, NNI_FAMILY), (2, 5, 0), (1, 8, 0)])
def test_synthetic_vs_literal_oracle(ctx, T, n, family):
    f = Forest.synthetic(T, n, seed=0x5EED0001 + T, family=family)
    o = orc.Forest([f.to_newick()])
    up, _, st = run_hashrf(ctx, f)
    assert np.array_equal(up.astype(np.int64), oracle_packed(o, "hashrf", "literal"))
    ost = {}
    o.hashrf(n, "closed", ost)
    assert st["n_unique"] == ost["n_unique"] and st["n_pairs"] == T * (T - 1) // 2


def test_config2_shape_vs_closed_form_oracle(ctx):
    # BASELINE config 2 family at a size the oracle finishes in seconds: 2500 x 100
    T, n = 2500, 100
    f = Forest.synthetic(T, n, seed=0x5EED0002)
    o = orc.Forest([f.to_newick()])
    small = orc.Forest([f.to_newick(0, 300)])
    assert np.array_equal(small.hashrf(n, "closed"), small.hashrf(n, "literal"))  # closed form is trusted only after this
    up, mask, st = run_hashrf(ctx, f, editdist=190)
    want = oracle_packed(o, "hashrf", "closed")
    assert np.array_equal(up.astype(np.int64), want)
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[: len(want)]
    assert np.array_equal(bits.astype(bool), want <= 190)
    assert 0 < bits.sum() < len(want)
    assert st["gpu_launches"] > 0 and st["variant_used"] == 1


def test_config3_shape_w4_vs_closed_form_oracle(ctx):
    # BASELINE config 3 family (200 taxa, W = 4 words per set) at a size the oracle finishes in seconds: the matrix through
    # the sparse kernel, through the tensor-core variant, and the threshold mask of both
    T, n = 2500, 200
    f = Forest.synthetic(T, n, seed=0x5EED0003)
    assert f.words == 4
    o = orc.Forest([f.to_newick()])
    small = orc.Forest([f.to_newick(0, 300)])
    assert np.array_equal(small.hashrf(n, "closed"), small.hashrf(n, "literal"))  # closed form is trusted only after this
    want = oracle_packed(o, "hashrf", "closed")
    bips, off = f.all_bips()
    for variant in (1, 2):
        up, mask, st = run_variant(ctx, bips, off, variant, editdist=388)
        assert st["variant_used"] == variant
        assert np.array_equal(up.astype(np.int64), want)
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[: len(want)]
        assert np.array_equal(bits.astype(bool), want <= 388)
        assert 0 < bits.sum() < len(want)


def test_multifurcating_nonuniform_bip_counts(ctx):
    rng = random.Random(11)
    text = [random_multifurcating_newicks(rng, 500, 40)]
    f, o = Forest.parse(text), orc.Forest(text)
    up, _, st = run_hashrf(ctx, f)
    assert st["min_bips"] < st["max_bips"]
    assert np.array_equal(up.astype(np.int64), oracle_packed(o, "hashrf", "literal"))


def test_quirk_q1_odd_taxa_on_device(ctx):
    # odd n: the tie rule makes rootings of one topology differ; device must agree
    rng = random.Random(5)
    text = [random_multifurcating_newicks(rng, 200, 9)]
    f, o = Forest.parse(text), orc.Forest(text)
    up, _, _ = run_hashrf(ctx, f)
    assert np.array_equal(up.astype(np.int64), oracle_packed(o, "hashrf", "literal"))


def test_duplicate_bips_within_a_tree_count_once(ctx):
    f = Forest.synthetic(50, 12, seed=3)
    bips, off = f.all_bips()
    want, _, _ = ctx.hashrf(bips, off)
    # repeat every entry of every tree twice: a tree's bipartitions are a Set
    rep = np.repeat(bips, 2, axis=0)
    got, _, st = ctx.hashrf(rep, off * 2)
    assert np.array_equal(got, want)


def test_editdist_mask_and_sharded_compute(ctx):
    T, n = 900, 30
    f = Forest.synthetic(T, n, seed=9, family=NNI_FAMILY, nni_max=4)
    bips, off = f.all_bips()
    full, mask, _ = ctx.hashrf(bips, off, editdist=4)
    P = T * (T - 1) // 2
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:P].astype(bool)
    assert np.array_equal(bits, full <= 4) and bits.any() and not bits.all()
    # staged API: 3 "ranks" each compute their aligned slice; slices tile [0, P) exactly
    ctx.hashrf_load(bips, off)
    got = np.empty(P, dtype=np.uint16)
    prev = 0
    for r in range(3):
        b, e = shard_range(T, r, 3)
        assert b == prev
        prev = e
        ctx.hashrf_compute(b, e, editdist=4)
        got[b:e] = ctx.fetch(b, e)
        m = np.unpackbits(ctx.fetch_mask(b, e).view(np.uint8), bitorder="little")[: e - b].astype(bool)
        assert np.array_equal(m, full[b:e] <= 4)
    assert prev == P and np.array_equal(got, full)
    ctx.hashrf_compute(0, P)
    rows = ctx.fetch_rows(10, 12)
    assert np.array_equal(rows, full[packed_index(T, 10, 11): packed_index(T, 12, 13)])


def _numpy_incidence(bips, off):
    """Independent restatement on the host: bipartition ids (exact, via np.unique on the raw words),
    per-id multiplicity, and a CSR id -> trees."""
    T = len(off) - 1
    keys = np.ascontiguousarray(bips).view([("", bips.dtype)] * bips.shape[1]).ravel()
    _, ids, counts = np.unique(keys, return_inverse=True, return_counts=True)
    tree_of = np.repeat(np.arange(T, dtype=np.int64), np.diff(off.astype(np.int64)))
    order = np.argsort(ids, kind="stable")
    starts = np.concatenate([[0], np.cumsum(counts)])
    return ids, counts, tree_of[order], starts


def test_large_T_multi_tile_rows_checksum_and_rows(ctx):
    # rows longer than one production tile (16384 columns): T = 40 000.  Too big for the oracle, so
    # check size-independent properties: the checksum identity
    #     sum_{a<b} d(a,b) = sum_{a<b} (|B_a| + |B_b|) - 2 * sum_lists m(m-1)/2
    # and complete rows recomputed on the host from an independent numpy incidence.
    T, n = 40_000, 100
    f = Forest.synthetic(T, n, seed=0x5EED0002)
    bips, off = f.all_bips()
    ids, counts, csr_trees, starts = _numpy_incidence(bips, off)
    nb = np.diff(off.astype(np.int64))
    st = ctx.hashrf_load(bips, off)
    P = T * (T - 1) // 2
    assert st["n_unique"] == len(counts)
    hits = int((counts.astype(np.int64) * (counts.astype(np.int64) - 1) // 2).sum())
    assert st["n_hits"] == hits
    st = ctx.hashrf_compute(0, P, editdist=2 * (n - 3) - 2)
    up = ctx.fetch(0, P)
    total = int(up.astype(np.uint64).sum())
    assert total == int(nb.sum()) * (T - 1) - 2 * hits
    mask = np.unpackbits(ctx.fetch_mask(0, P).view(np.uint8), bitorder="little")[:P].astype(bool)
    assert int(mask.sum()) == int((up <= 2 * (n - 3) - 2).sum())
    for a in [0, 1, 7, 16383, 16384, 23617, T - 16386, T - 300, T - 2]:
        shared = np.zeros(T, dtype=np.int64)
        for i in ids[int(off[a]):int(off[a + 1])]:
            np.add.at(shared, csr_trees[starts[i]:starts[i + 1]], 1)
        want = (nb[a] + nb[a + 1:] - 2 * shared[a + 1:]).astype(np.uint16)
        p0 = int(packed_index(T, a, a + 1))
        assert np.array_equal(up[p0:p0 + T - 1 - a], want), f"row {a}"
        assert np.array_equal(mask[p0:p0 + T - 1 - a], want <= 2 * (n - 3) - 2), f"mask row {a}"


def test_error_codes_not_aborts(ctx):
    fresh = RFContext(0)
    fresh.P = 0
    with pytest.raises(PhyBinRFError) as ei:
        fresh.hashrf_compute(0, 0)
    assert ei.value.code == -5  # PRF_E_STATE: compute before load
    fresh.close()
    f = Forest.synthetic(10, 8, seed=1)
    bips, off = f.all_bips()
    ctx.hashrf_load(bips, off)
    with pytest.raises(PhyBinRFError) as ei:
        ctx.hashrf_compute(1, 40)  # unaligned begin
    assert ei.value.code == -1
    bad = off.copy()
    bad[3] = bad[2] - 1
    with pytest.raises(PhyBinRFError):
        ctx.hashrf_load(bips, bad)


# ---- tolerant -------------------------------------------------------------------------------------

@pytest.mark.parametrize("seed", range(6))
def test_tolerant_random_shapes_vs_oracle(ctx, seed):
    rng = random.Random(100 + seed)
    n = rng.choice([3, 5, 8, 12, 30, 70])
    T = rng.randint(2, 60)
    text = [random_multifurcating_newicks(rng, T, n, p_missing=0.25, unary=0.1)]
    f, o = Forest.parse(text), orc.Forest(text)
    mat, _ = naiveDistMatrix(f, ctx)
    assert np.array_equal(mat.upper.astype(np.int64), oracle_packed(o, "naive")), text


def test_tolerant_config5_family_vs_oracle(ctx):
    # BASELINE config 5 family (150 taxa, each dropped w.p. 0.10) at an oracle-sized T
    T, n = 160, 150
    f = Forest.synthetic(T, n, seed=0x5EED0005, p_missing=0.10)
    o = orc.Forest([f.to_newick()])
    clades, off, leaf = f.raw_clades()
    up, mask, st = ctx.tolerant(clades, off, leaf, editdist=250)
    want = oracle_packed(o, "naive")
    assert np.array_equal(up.astype(np.int64), want)
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[: len(want)].astype(bool)
    assert np.array_equal(bits, want <= 250)


def test_tolerant_equals_hashrf_on_complete_taxa(ctx):
    f = Forest.synthetic(300, 24, seed=77)
    a = hashRF(24, f, ctx)
    b, _ = naiveDistMatrix(f, ctx)
    assert np.array_equal(a.upper, b.upper)


# ---- dense (tcgen05 int8 Gram) variant ---------------------------------------------------------

def run_variant(ctx, bips, off, variant, editdist=-1, p_begin=0, p_end=None):
    ctx.hashrf_load(bips, off)
    p_end = ctx.P if p_end is None else p_end
    st = ctx.hashrf_compute(p_begin, p_end, editdist=editdist, variant=variant)
    up = ctx.fetch(p_begin, p_end)
    mask = ctx.fetch_mask(p_begin, p_end) if editdist >= 0 else None
    return up, mask, st


@pytest.mark.parametrize("T,n,family,nni", [(300, 30, NNI_FAMILY, 4), (700, 64, NNI_FAMILY, 8), (1000, 100, NNI_FAMILY, 8),
                                            (513, 20, 0, 0), (2, 6, 0, 0), (257, 12, NNI_FAMILY, 2)])
def test_dense_variant_vs_literal_oracle(ctx, T, n, family, nni):
    f = Forest.synthetic(T, n, seed=0xD0 + T, family=family, nni_max=nni)
    o = orc.Forest([f.to_newick()])
    bips, off = f.all_bips()
    up, _, st = run_variant(ctx, bips, off, 2)
    assert st["variant_used"] == 2
    assert np.array_equal(up.astype(np.int64), oracle_packed(o, "hashrf", "literal"))


def test_dense_variant_multifurcating_mask_and_shards(ctx):
    rng = random.Random(21)
    text = [random_multifurcating_newicks(rng, 900, 24)]
    f, o = Forest.parse(text), orc.Forest(text)
    bips, off = f.all_bips()
    want = oracle_packed(o, "hashrf", "literal")
    T = len(f)
    P = T * (T - 1) // 2
    thr = int(np.median(want))
    up, mask, st = run_variant(ctx, bips, off, 2, editdist=thr)
    assert st["min_bips"] < st["max_bips"]
    assert np.array_equal(up.astype(np.int64), want)
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:P].astype(bool)
    assert np.array_equal(bits, want <= thr) and bits.any() and not bits.all()
    for r in range(3):
        b, e = shard_range(T, r, 3)
        got, m, _ = run_variant(ctx, bips, off, 2, editdist=thr, p_begin=b, p_end=e)
        assert np.array_equal(got.astype(np.int64), want[b:e])
        mb = np.unpackbits(m.view(np.uint8), bitorder="little")[: e - b].astype(bool)
        assert np.array_equal(mb, want[b:e] <= thr)


def test_majority_flip_keeps_distances_and_shrinks_the_join(ctx):
    # similar trees: the backbone's bipartitions are held by most trees and get stored as complement lists
    f = Forest.synthetic(1500, 60, seed=77, family=NNI_FAMILY, nni_max=3)
    o = orc.Forest([f.to_newick()])
    want = oracle_packed(o, "hashrf", "literal")
    bips, off = f.all_bips()
    flipped, _, st = run_variant(ctx, bips, off, 1)
    assert st["n_flipped"] > 0 and np.array_equal(flipped.astype(np.int64), want)
    ctx._ck(ctx._L.prf_debug_no_flip(ctx._h, 1))
    try:
        plain, _, st0 = run_variant(ctx, bips, off, 1)
    finally:
        ctx._ck(ctx._L.prf_debug_no_flip(ctx._h, 0))
    assert st0["n_flipped"] == 0 and np.array_equal(plain, flipped)
    assert st["n_hits"] * 4 < st0["n_hits"] and st["n_unique"] == st0["n_unique"]


def test_auto_variant_choice_follows_measured_density(ctx):
    f = Forest.synthetic(3000, 100, seed=5, family=NNI_FAMILY, nni_max=8)
    bips, off = f.all_bips()
    auto, _, st = run_variant(ctx, bips, off, 0)
    assert st["variant_used"] in (1, 2)   # decided by the cost model on the (flipped) incidence
    dense, _, st1 = run_variant(ctx, bips, off, 2)
    sparse, _, st2 = run_variant(ctx, bips, off, 1)
    assert st1["variant_used"] == 2 and st2["variant_used"] == 1
    assert np.array_equal(dense, sparse) and np.array_equal(auto, sparse)
    # with the majority flip switched off the same trees are a dense problem and AUTO takes the tensor cores
    ctx._ck(ctx._L.prf_debug_no_flip(ctx._h, 1))
    try:
        noflip, _, st3 = run_variant(ctx, bips, off, 0)
    finally:
        ctx._ck(ctx._L.prf_debug_no_flip(ctx._h, 0))
    assert st3["variant_used"] == 2 and np.array_equal(noflip, sparse)
    f = Forest.synthetic(3000, 100, seed=6)
    bips, off = f.all_bips()
    _, _, st = run_variant(ctx, bips, off, 0)
    assert st["variant_used"] == 1


# ---- the command-line host ----------------------------------------------------------------------

def test_cli_matches_reference_output_format(golden, tmp_path):
    import subprocess
    from phybin_b200 import build
    cli = build.build_cli()
    c = golden["t30_file"]
    tr = tmp_path / "mismatched.tr"
    tr.write_text(c["texts"][0])
    out = tmp_path / "out"
    r = subprocess.run([cli, "--tolerant", "--rfdist", "-o", str(out), str(tr)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert " Built matrix for dim 3" in r.stdout
    want = "Robinson-Foulds distance (matrix format):\n" + "-" * 41 + "\n0\n0 0\n1 0 0\n" + "-" * 41 + "\n"
    assert want in r.stdout and (out / "distance_matrix.txt").read_text() == want
    # default mode is --hashrf: the 9-taxon tree is filtered out (PhyBin.hs:187-194), 2 trees remain
    r = subprocess.run([cli, "--editdist=0", "--single", "-o", str(out), str(tr)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert " Using HashRF-style algorithm..." in r.stdout and " Built matrix for dim 2" in r.stdout
    assert (out / "distance_matrix.txt").read_text().splitlines()[2:4] == ["0", "1 0"]
    # --binary-matrix streams the packed triangle from the device
    t5 = tmp_path / "t5.tr"
    t5.write_text(golden["t5"]["texts"][0])
    r = subprocess.run([cli, "--binary-matrix", "-o", str(out), str(t5)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    m = DistanceMatrix.load_bin(out / "distance_matrix.bin")
    assert m.rows() == golden["t5"]["hashrf"]
    buf = io.StringIO()
    printDistMat(buf, m)
    assert buf.getvalue() == (out / "distance_matrix.txt").read_text()


@pytest.mark.parametrize("name,thr,linkage,host_bips", [("t4", 0, "--UPGMA", False), ("t5", 2, "--complete", False),
                                                        ("t3", 8, "--UPGMA", True), ("t3", 8, "--single", False)])
def test_cli_clusters_and_consensus_files(golden, tmp_path, name, thr, linkage, host_bips):
    """phybin --editdist on the reference's consensus fixtures: every tree of the file ends up in one cluster and
    cluster1_<size>_consensus.tr holds the reference's consensus bipartitions (for t4 the file is byte-identical
    to tests/t4_consensus/cluster1_16_consensus.tr)."""
    import subprocess
    from phybin_b200 import build
    cli = build.build_cli()
    c = golden[name]
    tr = tmp_path / "alltrees.tr"
    tr.write_text(c["texts"][0])
    out = tmp_path / "out"
    args = [cli, f"--editdist={thr}", linkage, "-o", str(out), str(tr)] + (["--host-bips"] if host_bips else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    T = c["hashrf_num_trees"]
    assert f"After flattening, cluster sizes are: [{T}]" in r.stdout and " Outcome: 1 clusters found, 1 non-singleton" in r.stdout
    assert (out / f"cluster1_{T}.txt").read_text().splitlines() == [f"alltrees_{i}" for i in range(T)]
    cons = (out / f"cluster1_{T}_consensus.tr").read_text()
    assert cons == c["consensus_newick"] + "\n"
    if name == "t4":
        assert cons.strip() == c["consensus_text"].strip()
    members = Forest.parse([(out / f"cluster1_{T}_alltrees.tr").read_text()])
    assert len(members) == T
    # dendrogram.txt (PhyBin.hs:238-243): the device's full agglomeration, shown as the reference's `show`; the same text
    # from the host restatement run on the golden matrix
    from phybin_b200 import host as phost
    rows = c["hashrf"]   # reference layout: row i holds d(i, j) for j < i
    upper = np.array([rows[b][a] for a in range(T) for b in range(a + 1, T)], dtype=np.uint16)
    _, ma, mb, md = phost.agglomerate(upper, T, {"--single": "single", "--complete": "complete", "--UPGMA": "UPGMA"}[linkage])
    assert (out / "dendrogram.txt").read_text() == phost.dendrogram_text([f"alltrees_{i}" for i in range(T)], ma, mb, md)
    assert " [finished] Wrote full dendrogram to file dendrogram.txt" in r.stdout
    # a tighter cut splits t5 / t3 into several clusters whose sizes add up
    if thr:
        r = subprocess.run([cli, "--editdist=0", linkage, "-o", str(tmp_path / "out0"), str(tr)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        line = [l for l in r.stdout.splitlines() if l.startswith("After flattening")][0]
        sizes = [int(x) for x in line.split("[")[1].rstrip("]").split(",")]
        assert sum(sizes) == T and sizes == sorted(sizes, reverse=True) and len(sizes) > 1


# ---- multi-GPU build (hash-partitioned dedup), emulated on one GPU ---------------------------------

def test_config4_shape_wide_sets_mask_vs_oracle(ctx):
    # BASELINE config 4 family (500 taxa, W = 8, --editdist mask) at an oracle-sized T.  A row has more
    # shared bipartitions than a walker keeps in registers, so the shared-memory cursor path runs too.
    T, n = 600, 500
    f = Forest.synthetic(T, n, seed=0x5EED0004)
    assert f.words == 8
    o = orc.Forest([f.to_newick()])
    bips, off = f.all_bips()
    want = oracle_packed(o, "hashrf", "closed")
    small = orc.Forest([f.to_newick(0, 60)])
    assert np.array_equal(small.hashrf(n, "closed"), small.hashrf(n, "literal"))
    thr = 2 * (n - 3) - 4
    for variant in (1, 2):
        up, mask, st = run_variant(ctx, bips, off, variant, editdist=thr)
        assert np.array_equal(up.astype(np.int64), want)
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[: len(want)].astype(bool)
        assert np.array_equal(bits, want <= thr)


def test_tolerant_wide_sets_vs_oracle(ctx):
    # 300 taxa with missing taxa: W = 5 words in the tolerant kernel
    T, n = 40, 300
    f = Forest.synthetic(T, n, seed=0x7015, p_missing=0.08)
    o = orc.Forest([f.to_newick()])
    clades, off, leaf = f.raw_clades()
    assert leaf.shape[1] == 5
    up, _, _ = ctx.tolerant(clades, off, leaf)
    assert np.array_equal(up.astype(np.int64), oracle_packed(o, "naive"))


# ---- flat clusters from the fused threshold mask --------------------------------------

@pytest.mark.parametrize("T,n,nni,thr", [(1200, 40, 3, 4), (700, 30, 6, 8), (300, 12, 0, 2), (5, 8, 1, 0)])
def test_mask_components_match_host_union_find(ctx, T, n, nni, thr):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    f = Forest.synthetic(T, n, seed=0xCC + T, family=NNI_FAMILY if nni else 0, nni_max=max(nni, 1))
    bips, off = f.all_bips()
    up, _, _ = run_variant(ctx, bips, off, 0, editdist=thr)
    labels, ncomp = ctx.mask_components()
    ii, jj = np.triu_indices(T, 1)              # row-major upper triangle == packed order
    close = up <= thr
    g = coo_matrix((np.ones(int(close.sum()), dtype=np.int8), (ii[close], jj[close])), shape=(T, T))
    want_n, want = connected_components(g, directed=False)
    assert ncomp == want_n
    # same partition, and every label is the smallest member of its component
    smallest = np.full(want_n, T, dtype=np.int64)
    np.minimum.at(smallest, want, np.arange(T))
    assert np.array_equal(labels.astype(np.int64), smallest[want])


def test_dedup_table_variants_agree(ctx):
    # the compact 4-byte-slot table (default below 2^25 entries) and the 8-byte-slot table give the same matrix
    f = Forest.synthetic(800, 70, seed=0xDED)
    o = orc.Forest([f.to_newick()])
    want = oracle_packed(o, "hashrf", "literal")
    bips, off = f.all_bips()
    a, _, st_a = run_variant(ctx, bips, off, 1)
    ctx._ck(ctx._L.prf_debug_wide_slots(ctx._h, 1))
    try:
        b, _, st_b = run_variant(ctx, bips, off, 1)
    finally:
        ctx._ck(ctx._L.prf_debug_wide_slots(ctx._h, 0))
    assert np.array_equal(a.astype(np.int64), want) and np.array_equal(a, b)
    assert st_a["n_unique"] == st_b["n_unique"] and st_a["n_hits"] == st_b["n_hits"]


# ---- allBips / raw clades extracted on the device -----------------------------------

def _assert_extraction_matches_host(ctx, f: Forest, W=None):
    W = W or f.words
    hb, ho = f.all_bips(W)
    info = ctx.allbips(f, EXTRACT_BIPS, W)
    db, do = ctx.allbips_fetch()
    assert info["nnz"] == len(hb) and np.array_equal(do, ho) and np.array_equal(db, hb)
    hc, hco, hl = f.raw_clades(W)
    info = ctx.allbips(f, EXTRACT_CLADES, W)
    dc, dco, dl = ctx.allbips_fetch()
    assert info["nnz"] == len(hc) and np.array_equal(dco, hco) and np.array_equal(dc, hc) and np.array_equal(dl, hl)


@pytest.mark.parametrize("T,n,family,p_missing", [(100, 10, 0, 0.0), (300, 64, 0, 0.0), (300, 65, 0, 0.0), (500, 100, 0, 0.0),
                                                  (400, 200, 0, 0.0), (257, 150, 0, 0.1), (300, 100, NNI_FAMILY, 0.0),
                                                  (64, 500, 0, 0.0), (33, 700, 0, 0.05), (1, 5, 0, 0.0)])
def test_device_allbips_equals_host_extraction(ctx, T, n, family, p_missing):
    f = Forest.synthetic(T, n, 0xA11B1 + T + n, family=family, p_missing=p_missing)
    _assert_extraction_matches_host(ctx, f)


@pytest.mark.parametrize("seed", range(4))
def test_device_allbips_odd_shapes_and_quirks(ctx, seed):
    """Multifurcations, unary chains, missing taxa (non-dense labels: Q2), odd sizes (Q1), trees of 1-3 leaves
    (where leaf singletons take part), wider-than-needed W."""
    rng = random.Random(1234 + seed)
    text = random_multifurcating_newicks(rng, 120, rng.choice([3, 4, 7, 13, 40, 70]), p_missing=0.3, unary=0.2)
    f = Forest.parse([text + "(x0);" + "x1;" + "((x0,x1));" + "((x2),(x0,x1),x3);"])
    _assert_extraction_matches_host(ctx, f)
    _assert_extraction_matches_host(ctx, f, W=f.words + 2)
    # and against the oracle's own allBips
    o = orc.Forest([text])
    g = Forest.parse([text])
    ctx.allbips(g, EXTRACT_BIPS)
    keys, off = ctx.allbips_fetch()
    for t in range(len(g)):
        dev = sorted(bits_to_labels(k) for k in keys[int(off[t]):int(off[t + 1])])
        assert dev == sorted(tuple(sorted(b)) for b in o.all_bips(t))


def test_device_allbips_feeds_the_matrix_without_host_bipartitions(ctx):
    f = Forest.synthetic(900, 64, 77)
    st = ctx.hashrf_load_trees(f)
    assert st["extract"]["gpu_launches"] >= 2 and st["nnz"] == st["extract"]["nnz"]
    ctx.hashrf_compute(editdist=3)
    got = ctx.fetch()
    bips, off = f.all_bips()
    want, _, _ = ctx.hashrf(bips, off)
    assert np.array_equal(got, want)
    g = Forest.synthetic(300, 40, 78, p_missing=0.15)
    ctx.tolerant_load_trees(g)
    ctx.tolerant_compute()
    got = ctx.fetch()
    want, _, _ = ctx.tolerant(*g.raw_clades())
    assert np.array_equal(got, want)
    assert np.array_equal(got, oracle_packed(orc.Forest([g.to_newick()]), "naive"))


def test_device_allbips_rejects_bad_input_with_codes(ctx):
    f = Forest.synthetic(10, 8, 5)
    label, parent, begin, leaves = (a.copy() for a in f.tree_arrays())
    L = _native.rf_lib()
    info = _native.prf_extract_info()

    def call(label, parent, begin, leaves, T, W, mode=0):
        return L.prf_allbips(ctx._h, label.ctypes.data, parent.ctypes.data, begin.ctypes.data, leaves.ctypes.data, T, W, 0,
                             mode, C.byref(info))

    assert call(label, parent, begin, leaves, 10, 1) == 0 and info.nnz == 50
    bad = parent.copy()
    bad[3] = 7  # a parent behind its child
    assert call(label, bad, begin, leaves, 10, 1) == -1 and b"pre-order" in L.prf_last_error(ctx._h)
    big = label.copy()
    big[label >= 0] += 64
    assert call(big, parent, begin, leaves, 10, 1) == -1
    assert call(label, parent, begin, leaves, 10, 0) == -6 and call(label, parent, begin, leaves, 10, 1, mode=7) == -1
    with pytest.raises(PhyBinRFError):  # a failed extraction leaves no result behind
        ctx.allbips_result()
    # shapes the kernel is not built for are refused with PRF_E_UNSUPPORTED, never silently rerouted:
    # a 9000-node unary chain above one leaf at W = 16 (9000 sets x 128 B do not fit a CTA's shared memory) ...
    m = 9000
    chain_label = np.full(m, -1, dtype=np.int32)
    chain_label[-1] = 0
    chain_parent = np.arange(-1, m - 1).astype(np.uint32)
    chain_begin = np.array([0, m], dtype=np.uint64)
    one = np.array([1], dtype=np.uint32)
    assert call(chain_label, chain_parent, chain_begin, one, 1, 16) == -6 and b"shared memory" in L.prf_last_error(ctx._h)
    assert call(chain_label, chain_parent, chain_begin, one, 1, 1) == 0 and info.nnz == 0  # the same chain fits at W = 1
    # ... and a tree of 20 000 nodes at any W
    m = 20000
    chain_label = np.full(m, -1, dtype=np.int32)
    chain_label[-1] = 0
    assert call(chain_label, np.arange(-1, m - 1).astype(np.uint32), np.array([0, m], dtype=np.uint64), one, 1, 1) == -6
    # a num_leaves that is not the tree's leaf count is an error, not a wrong result
    wrong = leaves.copy()
    wrong[2] += 1
    assert call(label, parent, begin, wrong, 10, 1) == -1
    # the widest sets the library takes (1000 taxa, W = 16), device and host extraction side by side
    wide = Forest.synthetic(3, 1000, 1)
    _assert_extraction_matches_host(ctx, wide)
    assert hashRF(1000, wide, ctx, extract="host").rows() == hashRF(1000, wide, ctx).rows()


# ---- strict consensus of every cluster on the device ---------------------------------------

def _cluster_sets(keys, off, labels):
    return {int(c): sorted(bits_to_labels(k) for k in keys[int(off[c]):int(off[c + 1])]) for c in np.unique(labels)}


@pytest.mark.parametrize("name", ["t3", "t4", "t5"])
def test_consensus_of_reference_fixture_clusters(golden, ctx, name):
    """The reference's consensus fixtures (Main.hs:470-510): one cluster holding every tree of the file."""
    c = golden[name]
    f = Forest.parse(c["texts"], c["name_prefix"])
    ctx.allbips(f, EXTRACT_BIPS)
    labels = np.zeros(len(f), dtype=np.uint32)
    keys, off, info = ctx.consensus(labels)
    assert info["n_clusters"] == 1 and int(off[1]) == info["n_out"] == len(c["consensus_bips"]) and int(off[-1]) == info["n_out"]
    assert sorted(bits_to_labels(k) for k in keys) == [tuple(b) for b in c["consensus_bips"]]
    assert f.consensus_newick(keys, c["consensus_num_taxa"]) == c["consensus_newick"]


@pytest.mark.parametrize("T,n,nni,thr,seed", [(600, 40, 3, 4, 1), (900, 100, 6, 8, 2), (300, 12, 0, 2, 3), (257, 130, 2, 2, 4)])
def test_consensus_per_mask_component_vs_oracle(ctx, T, n, nni, thr, seed):
    f = Forest.synthetic(T, n, 900 + seed, family=NNI_FAMILY if nni else 0, nni_max=max(nni, 1))
    o = orc.Forest([f.to_newick()])
    ctx.hashrf_load_trees(f)
    ctx.hashrf_compute(editdist=thr)
    labels, ncomp = ctx.mask_components()
    keys, off, info = ctx.consensus(labels)
    assert info["n_clusters"] == ncomp
    got = _cluster_sets(keys, off, labels)
    for c in np.unique(labels):
        assert got[int(c)] == o.consensus_bips(np.nonzero(labels == c)[0].tolist())
    non_reps = np.nonzero(labels != np.arange(T))[0]
    assert all(off[t] == off[t + 1] for t in non_reps)
    # same through host arrays, and with an arbitrary partition instead of mask components
    rng = np.random.default_rng(seed)
    group = rng.integers(0, 7, size=T)
    labels2 = np.array([np.nonzero(group == group[t])[0][0] for t in range(T)], dtype=np.uint32)
    bips, boff = f.all_bips()
    keys2, off2, _ = ctx.consensus(labels2, bips, boff)
    got2 = _cluster_sets(keys2, off2, labels2)
    for c in np.unique(labels2):
        assert got2[int(c)] == o.consensus_bips(np.nonzero(labels2 == c)[0].tolist())
        txt = f.consensus_newick(keys2[int(off2[c]):int(off2[c + 1])], n)
        back = orc.Forest([txt, f.to_newick(0, 1)])  # round trip: the printed tree has exactly the consensus bipartitions
        assert back.all_bips(0) == got2[int(c)]


def test_consensus_rejects_bad_labels(ctx):
    f = Forest.synthetic(20, 9, 3)
    bips, off = f.all_bips()
    ok = np.zeros(20, dtype=np.uint32)
    ok[10:] = 10
    keys, coff, info = ctx.consensus(ok, bips, off)
    assert info["n_clusters"] == 2 and np.count_nonzero(np.diff(coff.astype(np.int64))) <= 2
    bad = ok.copy()
    bad[3] = 5  # not the smallest id of its cluster
    with pytest.raises(PhyBinRFError) as e:
        ctx.consensus(bad, bips, off)
    assert e.value.code == -1
    with pytest.raises(PhyBinRFError):
        ctx._ck(_native.rf_lib().prf_consensus_fetch(ctx._h, None, None))


# ---- flat clusters for every linkage: device components + per-component host agglomeration ---

@pytest.mark.parametrize("T,n,nni,thr", [(150, 30, 4, 4), (120, 12, 0, 6), (90, 40, 8, 10)])
def test_flat_clusters_all_linkages_vs_naive_restatement(ctx, T, n, nni, thr):
    from oracle import pyref
    f = Forest.synthetic(T, n, 4242 + T, family=NNI_FAMILY if nni else 0, nni_max=max(nni, 1))
    ctx.hashrf_load_trees(f)
    ctx.hashrf_compute(editdist=thr)
    upper = ctx.fetch()
    d = lambda i, j: int(upper[packed_index(T, i, j)])
    ids = np.array(sorted(random.Random(T).sample(range(T), 17)), dtype=np.uint32)
    sub = ctx.fetch_submatrix(ids)
    assert sub.tolist() == [d(int(a), int(b)) for x, a in enumerate(ids) for b in ids[x + 1:]]
    comp, _ = ctx.mask_components()
    for linkage in ("single", "complete", "UPGMA"):
        got, k = ctx.flat_clusters(thr, linkage)
        assert got.tolist() == pyref.flat_clusters(d, T, linkage, thr) and k == len(set(got.tolist()))
        assert all(comp[t] == comp[got[t]] for t in range(T))  # every cluster lies inside one component
    with pytest.raises(PhyBinRFError):
        ctx.fetch_submatrix(np.array([5, 3], dtype=np.uint32))


def test_one_shot_from_trees_matches_staged_calls(ctx):
    f = Forest.synthetic(500, 70, 31)
    up, mask, st = ctx.hashrf_trees(f, editdist=5)
    want, wmask, _ = ctx.hashrf(*f.all_bips(), editdist=5)
    assert np.array_equal(up, want) and np.array_equal(mask, wmask)
    assert st["extract"]["nnz"] == st["nnz"] and st["gpu_launches"] > st["extract"]["gpu_launches"] > 0
    g = Forest.synthetic(200, 30, 32, p_missing=0.2)
    up, _, st = ctx.hashrf_trees(g, tolerant=True)
    want, _, _ = ctx.tolerant(*g.raw_clades())
    assert np.array_equal(up, want)


def test_flat_clusters_from_caller_owned_device_buffers(ctx):
    """The staged API lets the caller own the matrix and the mask (bench.py, multi-GPU ranks): the consumers read
    those buffers through their device pointers."""
    import torch
    T, thr = 400, 6
    f = Forest.synthetic(T, 30, 99, family=NNI_FAMILY, nni_max=4)
    ctx.hashrf_load_trees(f)
    ctx.hashrf_compute(editdist=thr)
    want = {lk: ctx.flat_clusters(thr, lk)[0] for lk in ("single", "complete", "UPGMA")}
    P = T * (T - 1) // 2
    d_out = torch.empty(P, dtype=torch.int16, device="cuda:0")
    d_mask = torch.empty((P + 31) // 32, dtype=torch.int32, device="cuda:0")
    torch.cuda.synchronize()
    ctx.hashrf_compute(0, P, thr, 0, d_out.data_ptr(), d_mask.data_ptr())
    for lk, w in want.items():
        got, _ = ctx.flat_clusters(thr, lk, d_mask.data_ptr(), d_out.data_ptr())
        assert np.array_equal(got, w)
    ids = np.array([3, 17, 44, 399], dtype=np.uint32)
    assert np.array_equal(ctx.fetch_submatrix(ids, d_out.data_ptr()), ctx.fetch()[[packed_index(T, int(a), int(b)) for x, a in enumerate(ids) for b in ids[x + 1:]]])

"""CPU tests (no GPU): the C++ host side against the oracle, the C-ABI library's exported
symbols, and loud failure without a device."""
import ctypes
import io
import os
import random
import re

import numpy as np
import pytest

from oracle import oracle as orc
from phybin_b200 import _native, build
from phybin_b200.host import Forest, NNI_FAMILY, bits_to_labels
from tests.util import random_multifurcating_newicks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_bips(f: Forest, t: int, bips, off):
    return sorted(bits_to_labels(b) for b in bips[int(off[t]):int(off[t + 1])])


@pytest.mark.parametrize("name", ["t30_literal", "t30_file", "t1", "t2", "t4", "t5", "t3"])
def test_host_allbips_matches_golden(golden, name):
    c = golden[name]
    f = Forest.parse(c["texts"], c["name_prefix"])
    assert f.labels == c["labels"]  # Parser.hs label order (quirk Q4)
    bips, off = f.all_bips()
    for t in range(len(f)):
        assert host_bips(f, t, bips, off) == sorted(tuple(b) for b in c["all_bips"][t])


def test_host_matches_oracle_on_random_shapes():
    rng = random.Random(7)
    for trial in range(30):
        n = rng.randint(2, 40)
        text = [random_multifurcating_newicks(rng, rng.randint(1, 6), n, p_missing=0.2, unary=0.1)]
        f, o = Forest.parse(text), orc.Forest(text)
        assert f.labels == o.labels
        bips, off = f.all_bips()
        for t in range(len(f)):
            assert f.num_leaves(t) == o.num_leaves(t)
            assert host_bips(f, t, bips, off) == sorted(o.all_bips(t)), (text, t)


def test_host_duplicate_labels_and_wide_sets():
    # duplicate taxa: numLeaves counts them twice (CoreTypes.hs:289) and so does the host
    text = ["((a,b),(a,c),(d,(e,f)));((a,b),(c,d),(e,f));"]
    f, o = Forest.parse(text), orc.Forest(text)
    assert f.has_duplicate_leaves()
    bips, off = f.all_bips()
    for t in range(2):
        assert host_bips(f, t, bips, off) == sorted(o.all_bips(t))
    # > 64 taxa: multi-word sets
    rng = random.Random(3)
    text = [random_multifurcating_newicks(rng, 4, 150)]
    f, o = Forest.parse(text), orc.Forest(text)
    assert f.words == 3
    bips, off = f.all_bips()
    for t in range(4):
        assert host_bips(f, t, bips, off) == sorted(o.all_bips(t))


@pytest.mark.parametrize("family,p_missing", [(0, 0.0), (NNI_FAMILY, 0.0), (0, 0.15)])
def test_synthetic_forest_roundtrips_through_newick(family, p_missing):
    f = Forest.synthetic(40, 37, seed=0x5EED0001, family=family, nni_max=6, p_missing=p_missing)
    text = f.to_newick()
    g, o = Forest.parse([text]), orc.Forest([text])
    assert f.labels == g.labels == o.labels  # generator numbers taxa exactly as the parser would
    b1, o1 = f.all_bips()
    b2, o2 = g.all_bips()
    assert np.array_equal(b1, b2) and np.array_equal(o1, o2)
    for t in range(len(f)):
        assert host_bips(f, t, b1, o1) == sorted(o.all_bips(t))
    if p_missing == 0:
        assert all(f.num_leaves(t) == 37 for t in range(40))
        assert all(int(o1[t + 1] - o1[t]) == 37 - 3 for t in range(40))  # binary: n - 3 bips


def test_synthetic_is_deterministic_and_seeded():
    a = Forest.synthetic(10, 20, seed=1).to_newick()
    assert a == Forest.synthetic(10, 20, seed=1).to_newick()
    assert a != Forest.synthetic(10, 20, seed=2).to_newick()


def test_raw_clades_shapes():
    text = ["((a,b),(c,(d,e)),f);(a,(b,c));"]
    f = Forest.parse(text)
    clades, off, leaf = f.raw_clades()
    lab = {n: i for i, n in enumerate(f.labels)}
    got0 = sorted(bits_to_labels(c) for c in clades[int(off[0]):int(off[1])])
    want0 = sorted(tuple(sorted(lab[x] for x in s)) for s in [("a", "b"), ("d", "e"), ("c", "d", "e")])
    assert got0 == want0
    assert bits_to_labels(leaf[1]) == tuple(sorted(lab[x] for x in "abc"))


def test_filter_num_leaves_is_the_hashrf_input_filter(golden):
    c = golden["t30_literal"]
    f = Forest.parse(c["texts"])
    assert len(f) == 3 and f.filter_num_leaves(10) == 2  # PhyBin.hs:187-194 drops the 9-taxon tree


def test_print_dist_mat_text():
    L = _native.host_lib()
    T = 3
    upper = np.array([0, 1, 0], dtype=np.uint16)  # d(0,1)=0 d(0,2)=1 d(1,2)=0  ==  [[],[0],[1,0]]
    p = L.phb_print_dist_mat(upper.ctypes.data, T)
    s = ctypes.string_at(p).decode()
    L.phb_free_buf(p)
    assert s == orc.print_dist_mat(np.array([0, 1, 0]), 3)
    assert s.splitlines()[2:5] == ["0", "0 0", "1 0 0"]


def test_parse_errors_are_reported_not_fatal():
    with pytest.raises(ValueError):
        Forest.parse(["((a,b),c"])
    with pytest.raises(ValueError):
        orc.Forest(["((a,b),c"])


def test_rf_library_exports_every_declared_symbol():
    """include/phybin_rf.h <-> libphybin_rf.so: every declared entry point is exported."""
    build.build_cuda()
    hdr = open(os.path.join(ROOT, "include", "phybin_rf.h")).read()
    declared = sorted(set(re.findall(r"PRF_API\s+[\w\s\*]+?\b(prf_\w+)\s*\(", hdr)))
    assert declared == sorted(_native.RF_SYMBOLS)
    lib = ctypes.CDLL(build.RF_SO)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_rf_helpers_without_a_device():
    L = _native.rf_lib()
    T = 1000
    assert L.prf_packed_index(T, 0, 1) == 0
    assert L.prf_packed_index(T, 0, T - 1) == T - 2
    assert L.prf_packed_index(T, T - 2, T - 1) == T * (T - 1) // 2 - 1
    assert L.prf_packed_index(T, 5, 5) == 2**64 - 1
    align = L.prf_shard_align()
    P = 100000 * 99999 // 2
    prev = 0
    for r in range(8):
        b, e = ctypes.c_uint64(), ctypes.c_uint64()
        assert L.prf_shard_range(100000, r, 8, ctypes.byref(b), ctypes.byref(e)) == 0
        assert b.value == prev and b.value % align == 0
        prev = e.value
    assert prev == P


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phybin_b200.rfdistance import PhyBinRFError, RFContext
    with pytest.raises(PhyBinRFError) as ei:
        RFContext(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


@pytest.mark.parametrize("T,tile", [(2, 256), (3, 256), (50, 256), (700, 256), (5000, 16384), (100000, 16384)])
def test_sparse_kernel_work_plan_tiles_every_shard_exactly(T, tile):
    """Host logic of the matrix kernel: the work items (long rows / groups of short rows) of every shard
    of every world size cover the shard's packed range exactly once."""
    L = _native.rf_lib()
    P = T * (T - 1) // 2
    row_start = lambda a: a * T - a * (a + 1) // 2  # noqa: E731
    buf = np.zeros((T + 8, 4), dtype=np.uint32)
    for world in (1, 3, 8):
        for rank in range(world):
            b, e = ctypes.c_uint64(), ctypes.c_uint64()
            assert L.prf_shard_range(T, rank, world, ctypes.byref(b), ctypes.byref(e)) == 0
            n = L.prf_debug_row_plan(T, b.value, e.value, tile, buf.ctypes.data, len(buf))
            assert n >= 0
            spans = []
            for a0, a1, cb, ce in buf[:n].tolist():
                if a0 == a1:      # one row, columns [cb, ce)
                    assert a0 + 1 <= cb < ce <= T
                    spans.append((row_start(a0) + cb - a0 - 1, row_start(a0) + ce - a0 - 1))
                else:             # whole rows a0..a1 in one tile
                    lo, hi = row_start(a0), row_start(a1 + 1)
                    assert a1 > a0 and hi - lo <= tile and a1 - a0 + 1 <= 256
                    spans.append((lo, hi))
            spans.sort()
            pos = b.value
            for lo, hi in spans:
                assert lo == pos and hi > lo
                pos = hi
            assert pos == e.value
    assert P == row_start(T - 1)


def test_host_library_exports_every_declared_symbol():
    """include/phybin_host.h <-> libphybin_host.so."""
    build.build_host()
    hdr = open(os.path.join(ROOT, "include", "phybin_host.h")).read()
    declared = sorted(set(re.findall(r"PHB_API\s+[\w\s\*]+?\b(phb_\w+)\s*\(", hdr)))
    assert len(declared) >= 15
    lib = ctypes.CDLL(build.HOST_SO)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_print_dist_mat_streams_the_same_text_to_a_file(tmp_path):
    L = _native.host_lib()
    rng = np.random.default_rng(5)
    T = 37
    upper = rng.integers(0, 70, size=T * (T - 1) // 2, dtype=np.uint16)
    p = L.phb_print_dist_mat(upper.ctypes.data, T)
    text = ctypes.string_at(p).decode()
    L.phb_free_buf(p)
    path = tmp_path / "distance_matrix.txt"
    assert L.phb_write_dist_mat(str(path).encode(), upper.ctypes.data, T) == 0
    ii, jj = np.tril_indices(T, -1)  # the reference's layout: row i holds d(i, 0..i-1)
    lower = upper[(jj * T - jj * (jj + 1) // 2 + (ii - jj - 1))].astype(np.int64)
    assert path.read_text() == text == orc.print_dist_mat(lower, T)


# ---- clustering consumer: the host step that follows prf_mask_components ---------------------

def _canon(lbl):
    first = {}
    return [first.setdefault(int(l), i) for i, l in enumerate(lbl)]


@pytest.mark.parametrize("linkage,method", [("single", "single"), ("complete", "complete"), ("UPGMA", "average")])
def test_cluster_component_matches_scipy_on_tie_free_distances(linkage, method):
    from scipy.cluster.hierarchy import fcluster, linkage as scipy_linkage
    from scipy.spatial.distance import pdist
    from phybin_b200.host import cluster_component
    rng = np.random.default_rng(11)
    for m in (2, 3, 10, 60, 250):
        d = pdist(rng.random((m, 3)))  # condensed form == packed row-major upper triangle
        for thr in (0.05, 0.2, 0.5, 2.0):
            got = cluster_component(d, m, linkage, thr)
            want = fcluster(scipy_linkage(d, method=method), t=thr, criterion="distance")
            assert _canon(got) == _canon(want), (m, thr)


@pytest.mark.parametrize("linkage", ["single", "complete", "UPGMA"])
def test_cluster_component_tie_order_matches_naive_restatement(linkage):
    """Integer distances with many ties: the nearest-neighbour-cached agglomeration picks exactly the pairs the
    naive distance-matrix algorithm (oracle/pyref.flat_clusters) picks."""
    from oracle import pyref
    from phybin_b200.host import cluster_component
    rng = np.random.default_rng(5)
    for m in (1, 2, 5, 17, 40):
        for _ in range(4):
            A = np.triu(rng.integers(0, 6, size=(m, m)), 1)
            sub = np.array([A[i, j] for i in range(m) for j in range(i + 1, m)], dtype=np.uint16)
            for thr in (0, 1, 2, 4):
                assert cluster_component(sub, m, linkage, thr).tolist() == pyref.flat_clusters(lambda i, j: A[i, j], m, linkage, thr)


def test_binary_matrix_round_trip(tmp_path):
    """the packed binary form of the matrix (header + uint16 upper triangle), written in pieces."""
    from phybin_b200.rfdistance import DistanceMatrix
    rng = np.random.default_rng(3)
    for T in (0, 1, 2, 7, 300):
        up = rng.integers(0, 995, size=T * (T - 1) // 2 if T else 0).astype(np.uint16)
        path = tmp_path / f"m{T}.bin"
        DistanceMatrix(up, T).save_bin(path, chunk=1000)
        assert path.stat().st_size == 32 + 2 * len(up)
        back = DistanceMatrix.load_bin(path)
        assert back.T == T and np.array_equal(back.upper, up)
    (tmp_path / "bad.bin").write_bytes(b"PHYBINRX" + bytes(24))
    with pytest.raises(OSError):
        DistanceMatrix.load_bin(tmp_path / "bad.bin")


def test_bips_to_tree_order_matches_restatement_on_wide_sets():
    """phb_consensus_newick against oracle/pyref.bips_to_tree on random trees of up to 260 taxa (multi-word sets,
    many equal-sized bipartitions: exercises the IntSet tie order of the size sort, RFDistance.hs:413-414)."""
    import random
    from oracle import pyref
    from tests.util import random_multifurcating_newicks
    rng = random.Random(5)
    for n in (4, 5, 9, 40, 70, 130, 260):
        text = random_multifurcating_newicks(rng, 8, n)
        names, trees = pyref.parse_newicks([text])
        f = Forest.parse([text])
        assert f.labels == names
        bips, off = f.all_bips()
        for t, tr in enumerate(trees):
            want = pyref.to_newick(names, pyref.bips_to_tree(n, sorted(pyref.all_bips(tr))))
            rows = bips[int(off[t]):int(off[t + 1])]
            assert f.consensus_newick(rows[::-1], n) == want  # any input order
    # a set that no subtree lies inside is the reference's "No match for bip" error
    f = Forest.parse(["(a,b,c,d);"])
    with pytest.raises(ValueError):
        f.consensus_newick(np.zeros((1, 1), dtype=np.uint64), 4)

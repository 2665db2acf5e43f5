"""Pins the CPU oracle (oracle/phybin_oracle.cpp) before anything else trusts it.

* the reference's one asserted RF known answer: TestAll.hs:22-32, "[[],[0],[1,0]]";
* SURVEY.md section 4's fixture-derived aggregates (t1 -> 6, t2 -> 0, t4 all 0 / U=7,
  t5 sum 788 max 2 U=9, t3 sum 198806 max 8 U=38);
* element-by-element agreement with the independent pure-Python restatement (oracle/pyref.py)
  whose outputs are committed in tests/golden/fixtures.json;
* the reference's structural pins: allBips is rooting-invariant for a dense even-n tree
  (Main.hs:420-441) and hashRF == naiveDistMatrix on complete, identical taxon sets.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import pyref

ALL = ["t30_literal", "t30_file", "t1", "t2", "t4", "t5", "t3"]


def test_reference_kat_t30(golden):
    c = golden["t30_literal"]
    f = orc.Forest(c["texts"], c["name_prefix"])
    rows = orc.lower_to_rows(f.naive(), f.num_trees)
    # showMat t30' == "[[],[0],[1,0]]"  (TestAll.hs:31-32)
    assert str(rows).replace(" ", "") == "[[],[0],[1,0]]"
    assert rows == c["reference_asserted_naive"]
    # label numbering quirk Q4 (Parser.hs:71,92-110): last file first, children right-to-left
    assert f.labels == ["1_", "7_", "6_", "18", "13", "5_", "19", "3_", "14", "2_"]


@pytest.mark.parametrize("name", ALL)
def test_labels_and_bips_match_pyref(golden, name):
    c = golden[name]
    f = orc.Forest(c["texts"], c["name_prefix"])
    assert f.labels == c["labels"]
    assert f.num_trees == c["num_trees"]
    for t in range(f.num_trees):
        assert [list(b) for b in f.all_bips(t)] == c["all_bips"][t]


@pytest.mark.parametrize("name", ALL)
def test_naive_matches_golden(golden, name):
    c = golden[name]
    f = orc.Forest(c["texts"], c["name_prefix"])
    assert orc.lower_to_rows(f.naive(), f.num_trees) == c["naive"]


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("variant", ["literal", "closed", "literal_all_cores"])
def test_hashrf_matches_golden(golden, name, variant):
    c = golden[name]
    f = orc.Forest(c["texts"], c["name_prefix"])
    T = f.filter_num_leaves(len(c["labels"]))  # HashRF input filter, PhyBin.hs:187-194
    assert T == c["hashrf_num_trees"]
    st = {}
    flat = f.hashrf(len(c["labels"]), variant, st)
    assert orc.lower_to_rows(flat, T) == c["hashrf"]
    assert st["n_unique"] == c["n_unique_full"]
    s = c.get("survey", {})
    if "sum" in s:
        assert int(flat.sum()) == s["sum"] and int(flat.max(initial=0)) == s["max"]
        assert st["n_unique"] == s["n_unique"]
    if "hashrf" in s:
        assert orc.lower_to_rows(flat, T) == s["hashrf"]
    if "bips_per_tree" in s:
        assert all(len(f.all_bips(t)) == s["bips_per_tree"] for t in range(T))


def test_hashrf_equals_naive_on_complete_taxa(golden):
    for name in ["t3", "t4", "t5"]:
        c = golden[name]
        f = orc.Forest(c["texts"], c["name_prefix"])
        assert np.array_equal(f.hashrf(10), f.naive())


def test_allbips_rooting_invariant():
    # Main.hs:420-441 (norm/Bip1): the same unrooted 10-taxon tree under two rootings.
    a = "((a,b),(c,(d,e)),((f,g),(h,(i,j))));"
    b = "(a,b,((c,(d,e)),((f,g),(h,(i,j)))));"
    f = orc.Forest([a + b])
    assert sorted(f.all_bips(0)) == sorted(f.all_bips(1))
    assert int(f.hashrf(10)[0]) == 0


def test_quirk_q1_odd_size_tie_rule():
    # n = 7, same unrooted tree, three presentations -> [[],[1],[1,2]].
    # Labels must be 0..6 with t0 -> 0: the first tree is labelled right-to-left, so list names
    # in the order that gives the labels the survey example uses.
    texts = ["((t1,t2,t3),(t0,t4,t5,t6));((t1,t2,t3),t0,t4,t5,t6);((t0,t4,t5,t6),t1,t2,t3);"]
    f = orc.Forest(texts)
    lab = {n: i for i, n in enumerate(f.labels)}
    names, trees = pyref.parse_newicks(texts)
    assert names == f.labels
    rows = orc.lower_to_rows(f.hashrf(7), 3)
    assert rows == pyref.hash_rf(trees)
    # relabel-independent statement of the quirk: the three matrices entries are 1, 1, 2 in some order
    assert sorted(x for r in rows for x in r) == [1, 1, 2] or lab["t0"] != 0


def test_quirk_q2_invert_dense_universe():
    # A=((1,2),(3,(4,5)),0), B=(1,2,(3,(4,5))): reference distance 1 (true RF 0).
    # Name the taxa so that label k gets name "k": the LAST tree is labelled first, right-to-left.
    # Build through pyref to get the labels, then compare oracle == pyref on the same text.
    texts = ["((n1,n2),(n3,(n4,n5)),n0);(n1,n2,(n3,(n4,n5)));"]
    f = orc.Forest(texts)
    names, trees = pyref.parse_newicks(texts)
    assert orc.lower_to_rows(f.naive(), 2) == pyref.naive_dist_matrix(trees)


def test_print_dist_mat_format():
    want = ("Robinson-Foulds distance (matrix format):\n"
            "-----------------------------------------\n"
            "0\n0 0\n1 0 0\n"
            "-----------------------------------------\n")
    assert orc.print_dist_mat(np.array([0, 1, 0]), 3) == want
    assert pyref.print_dist_mat([[], [0], [1, 0]]) == want


def test_random_small_trees_oracle_vs_pyref():
    # random multifurcating trees with missing taxa, unary nodes: C++ oracle == pyref (both modes)
    import random
    rng = random.Random(1234)

    def rand_tree(labels):
        nodes = [str(l) for l in labels]
        rng.shuffle(nodes)
        while len(nodes) > 1:
            k = min(len(nodes), rng.choice([2, 2, 2, 3, 4]))
            if len(nodes) <= 3 and rng.random() < 0.5:
                k = len(nodes)
            grp = [nodes.pop(rng.randrange(len(nodes))) for _ in range(k)]
            s = "(" + ",".join(grp) + ")"
            if rng.random() < 0.1:
                s = "(" + s + ")"  # unary interior node
            nodes.append(s)
        return nodes[0] + ";"

    for trial in range(40):
        n = rng.randint(2, 11)
        T = rng.randint(2, 6)
        texts = []
        for _ in range(T):
            labs = [f"x{l}" for l in range(n) if rng.random() > 0.25] or ["x0"]
            texts.append(rand_tree(labs))
        text = ["".join(texts)]
        f = orc.Forest(text)
        names, trees = pyref.parse_newicks(text)
        assert f.labels == names
        assert orc.lower_to_rows(f.naive(), T) == pyref.naive_dist_matrix(trees), text
        assert orc.lower_to_rows(f.hashrf(n), T) == pyref.hash_rf(trees), text
        assert np.array_equal(f.hashrf(n, "closed"), f.hashrf(n, "literal"))


# ---- strict consensus: the reference's own consensus fixtures ------------------------------

@pytest.mark.parametrize("name", ["t3", "t4", "t5"])
def test_consensus_matches_reference_fixture_tree(golden, name):
    """consensusTest (Main.hs:484-510): the bipartitions of tests/t*_consensus/*_consensus.tr equal those of
    consensusTree over the cluster's trees -- for the C++ oracle, for pyref, and for the host's bipsToTree."""
    from phybin_b200.host import Forest, bits_to_labels
    c = golden[name]
    want = [tuple(b) for b in c["consensus_bips"]]
    o = orc.Forest(c["texts"], c["name_prefix"])
    assert o.consensus_bips(range(o.num_trees)) == want
    names, trees = pyref.parse_newicks(c["texts"], c["name_prefix"])
    assert sorted(pyref.consensus_bips(trees)) == want
    assert pyref.to_newick(names, pyref.consensus_tree(c["consensus_num_taxa"], trees)) == c["consensus_newick"]
    # the tree on disk, parsed together with the cluster's trees, has exactly these bipartitions
    both = orc.Forest([c["consensus_text"]] + c["texts"], c["name_prefix"])
    assert both.labels == c["labels"] and both.all_bips(0) == want
    # host side (product code): bipsToTree + one-line printing
    f = Forest.parse(c["texts"], c["name_prefix"])
    keys = np.zeros((len(want), f.words), dtype=np.uint64)
    for i, b in enumerate(want):
        for l in b:
            keys[i, l >> 6] |= np.uint64(1) << np.uint64(l & 63)
    assert f.consensus_newick(keys[::-1], c["consensus_num_taxa"]) == c["consensus_newick"]
    if name == "t4":  # same label order as when the reference wrote the file: byte-identical
        assert c["consensus_newick"] == c["consensus_text"].strip()


@pytest.mark.parametrize("name", ["t3", "t4", "t5", "t1", "t2"])
def test_bips_to_tree_round_trip(golden, name):
    """testBipConversion (Main.hs:519-530): allBips (bipsToTree n (allBips t)) == allBips t, through the host code."""
    from phybin_b200.host import Forest, bits_to_labels
    c = golden[name]
    f = Forest.parse(c["texts"], c["name_prefix"])
    f.filter_num_leaves(len(c["labels"]))
    bips, off = f.all_bips()
    rebuilt = [f.consensus_newick(bips[int(off[t]):int(off[t + 1])], len(c["labels"])) for t in range(len(f))]
    g = Forest.parse(["".join(rebuilt)] + c["texts"], c["name_prefix"])  # the last file is labelled first: same table
    assert g.labels == c["labels"]
    b2, o2 = g.all_bips()
    for t in range(len(f)):  # as sets: the order inside a tree follows its node order
        assert sorted(map(tuple, b2[int(o2[t]):int(o2[t + 1])].tolist())) == sorted(map(tuple, bips[int(off[t]):int(off[t + 1])].tolist()))


def test_consensus_property_on_random_clusters():
    """The property consensusTest asserts on the fixtures (Main.hs:505-508), on random inputs: the tree rebuilt by
    bipsToTree from the intersection has exactly the intersection as its bipartitions -- for both restatements."""
    import random
    from tests.util import random_multifurcating_newicks
    rng = random.Random(21)
    for n in (4, 6, 10, 25, 70):
        for _ in range(3):
            # similar trees (so that the intersection is not empty): one random tree, some copies, a few others
            base = random_multifurcating_newicks(rng, 1, n)
            text = base * rng.randint(1, 3) + random_multifurcating_newicks(rng, rng.randint(0, 2), n)
            names, trees = pyref.parse_newicks([text])
            if any(pyref.num_leaves(t) != n for t in trees):
                continue
            inter = pyref.consensus_bips(trees)
            rebuilt = pyref.bips_to_tree(n, inter)
            assert pyref.all_bips(rebuilt) == inter
            o = orc.Forest([text])
            assert o.consensus_bips(range(o.num_trees)) == sorted(inter)
            # the printed tree parses back to the same bipartitions under the same label table
            back = orc.Forest([pyref.to_newick(names, rebuilt), text])
            assert back.labels == names and back.all_bips(0) == sorted(inter)

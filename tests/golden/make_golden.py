"""Generates tests/golden/fixtures.json.  Run HERE (needs /root/reference); the GPU box only reads
the committed JSON.

Inputs  : the reference's own Newick fixtures (tests/t1..t5, t30_mismatched) and the three literal
          trees of its one asserted RF known-answer test (TestAll.hs:22-32).
Expected: (a) the reference-asserted answer for t30 ("[[],[0],[1,0]]", TestAll.hs:31-32);
          (b) SURVEY.md section 4's fixture-derived aggregates (sum / max / #unique bips);
          (d) for t3/t4/t5 the reference's consensus fixtures (tests/t*_consensus/*_consensus.tr): the
              bipartitions of the tree ON DISK, which the reference's own test asserts equal to those of
              consensusTree over the cluster's trees (Main.hs:470-510);
          (c) full matrices from oracle/pyref.py, the independent pure-Python restatement, so the
              C++ oracle and the CUDA path are compared element by element.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

REF = "/root/reference/tests"


def rd(p):
    with open(os.path.join(REF, p)) as f:
        return f.read()


CASES = [
    # name, files (command-line order), name_prefix (-p), survey aggregates
    dict(name="t30_literal", texts=["(2_, (((14, 3_), (19, (5_, 13))), (18, (6_, 7_))), 1_);",
                                    "(2_, (((14, 3_), (19, (5_, 13))), (6_, 7_)), 1_);",
                                    "(2_, (((14, 3_), (19, (5_, 13))), (18, 6_, 7_)), 1_);"],
         name_prefix=-1, source="TestAll.hs:22-26 (three (name, string) pairs)",
         reference_asserted_naive=[[], [0], [1, 0]]),
    dict(name="t30_file", files=["t30_mismatched/mismatched.tr"], name_prefix=-1,
         source="tests/t30_mismatched/mismatched.tr", survey=dict(naive=[[], [0], [1, 0]])),
    dict(name="t1", files=["t1/112.tr", "t1/13.tr"], name_prefix=2, source="tests/t1 with -p 2",
         survey=dict(hashrf=[[], [6]], naive=[[], [6]])),
    dict(name="t2", files=["t2/116.tr", "t2/233.tr"], name_prefix=2, source="tests/t2 with -p 2",
         survey=dict(hashrf=[[], [0]])),
    dict(name="t4", files=["t4_consensus/cluster1_16_alltrees.tr"], consensus_file="t4_consensus/cluster1_16_consensus.tr", name_prefix=-1,
         source="tests/t4_consensus", survey=dict(sum=0, max=0, n_unique=7, bips_per_tree=7)),
    dict(name="t5", files=["t5_consensus/cluster1_35_alltrees.tr"], consensus_file="t5_consensus/cluster1_35_consensus.tr", name_prefix=-1,
         source="tests/t5_consensus", survey=dict(sum=788, max=2, n_unique=9)),
    dict(name="t3", files=["t3_consensus/cluster1_284_alltrees.tr"], consensus_file="t3_consensus/cluster1_284_consensus.tr", name_prefix=-1,
         source="tests/t3_consensus", survey=dict(sum=198806, max=8, n_unique=38)),
]

out = []
for c in CASES:
    texts = c.pop("texts", None) or [rd(p) for p in c["files"]]
    names, trees = pyref.parse_newicks(texts, c["name_prefix"])
    n = len(names)
    full = [t for t in trees if pyref.num_leaves(t) == n]
    rec = dict(c)
    rec["texts"] = texts
    rec["labels"] = names
    rec["num_trees"] = len(trees)
    rec["naive"] = pyref.naive_dist_matrix(trees)
    # hashRF sees only trees with numLeaves == #labels (PhyBin.hs:187-194)
    rec["hashrf_num_trees"] = len(full)
    rec["hashrf"] = pyref.hash_rf(full)
    bips = [sorted(pyref.all_bips(t)) for t in trees]
    rec["all_bips"] = [[list(b) for b in bs] for bs in bips]
    rec["n_unique_full"] = len({b for t in full for b in pyref.all_bips(t)})
    if "consensus_file" in c:
        # consensusTest (Main.hs:484-510): parse [consensus, alltrees] together (the last file is labelled first,
        # so the label table is the one of alltrees alone); the tree on disk must have exactly the bipartitions
        # of consensusTree num_taxa trees.
        ctext = rd(c["consensus_file"])
        names2, both = pyref.parse_newicks([ctext] + texts, c["name_prefix"])
        assert names2 == names
        ctree = both[0]
        on_disk = pyref.all_bips(ctree)
        recomputed = pyref.consensus_tree(pyref.num_leaves(ctree), trees)
        assert on_disk == pyref.all_bips(recomputed) == pyref.consensus_bips(trees), c["name"]
        rec["consensus_text"] = ctext
        rec["consensus_num_taxa"] = pyref.num_leaves(ctree)
        rec["consensus_bips"] = [list(b) for b in sorted(on_disk)]
        rec["consensus_newick"] = pyref.to_newick(names, recomputed)
    out.append(rec)
    flat = [x for r in rec["hashrf"] for x in r]
    print(c["name"], "T=%d labels=%d hashrf: sum=%d max=%d U=%d" % (
        len(trees), n, sum(flat), max(flat or [0]), rec["n_unique_full"]))

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))

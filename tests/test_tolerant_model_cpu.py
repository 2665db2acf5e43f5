"""The tolerant kernel's per-pair recipe (tests/tolerant_model.py mirrors prf_tolerant.cuh) against the oracle's literal
naiveDistMatrix (RFDistance.hs:168-198) on the arrays the host hands to the device -- including forests whose trees
REPEAT a leaf label (numLeaves counts them twice, CoreTypes.hs:289-291)."""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from phybin_b200.host import Forest
from tests.tolerant_model import pair_distance, parents
from tests.util import oracle_packed, random_multifurcating_newicks


def dup_newicks(rng: random.Random, T: int, n: int, p_missing: float, p_dup: float, unary: float = 0.1) -> str:
    """Random trees in which a leaf label may occur several times."""
    out = []
    for _ in range(T):
        labs = [f"x{l}" for l in range(n) if rng.random() >= p_missing] or ["x0"]
        labs += [rng.choice(labs) for _ in range(len(labs)) if rng.random() < p_dup]
        nodes = labs[:]
        rng.shuffle(nodes)
        while len(nodes) > 1:
            k = min(len(nodes), rng.choice([2, 2, 3, 4]))
            grp = [nodes.pop(rng.randrange(len(nodes))) for _ in range(k)]
            s = "(" + ",".join(grp) + ")"
            if rng.random() < unary:
                s = "(" + s + ")"
            nodes.append(s)
        out.append(nodes[0] + ";")
    return "".join(out)


def _check(text, dups):
    f, o = Forest.parse([text]), orc.Forest([text])
    T = len(f)
    want = oracle_packed(o, "naive")
    if dups:
        clades, outs, off, leaf, doff, dlab, dext = f.raw_clades_ex()
        got = [pair_distance(a, b, clades, off, leaf, outsets=outs, dup=(doff, dlab, dext)) for a in range(T) for b in range(a + 1, T)]
    else:
        clades, off, leaf = f.raw_clades()
        par = parents(clades, off)
        got = [pair_distance(a, b, clades, off, leaf, par=par) for a in range(T) for b in range(a + 1, T)]
    assert got == want.tolist()


@pytest.mark.parametrize("seed", range(6))
def test_parent_skip_recipe_vs_oracle(seed):
    rng = random.Random(seed)
    n = rng.choice([3, 4, 5, 8, 20, 70])
    _check(random_multifurcating_newicks(rng, 14, n, p_missing=rng.choice([0.0, 0.2, 0.5]), unary=rng.choice([0.0, 0.3])), False)


@pytest.mark.parametrize("seed", range(10))
def test_repeated_label_recipe_vs_oracle(seed):
    rng = random.Random(100 + seed)
    n = rng.choice([2, 3, 4, 5, 9, 30, 66])
    text = dup_newicks(rng, 12, n, p_missing=rng.choice([0.0, 0.15, 0.4]), p_dup=rng.choice([0.05, 0.3]))
    _check(text, True)


def test_repeated_label_form_equals_distinct_form_without_repeats():
    rng = random.Random(5)
    text = random_multifurcating_newicks(rng, 12, 25, p_missing=0.2, unary=0.2)
    f = Forest.parse([text])
    assert not f.has_duplicate_leaves()
    _check(text, True)


@pytest.mark.parametrize("seed", range(5))
def test_both_oracles_agree_on_repeated_labels(seed):
    """The C++ oracle and the independent pure-Python restatement (oracle/pyref.py) on forests that repeat leaf labels:
    numLeaves counts the repeats (CoreTypes.hs:289-291), label sets do not."""
    from oracle import pyref
    rng = random.Random(900 + seed)
    text = dup_newicks(rng, 9, rng.choice([3, 5, 12, 40]), p_missing=rng.choice([0.0, 0.3]), p_dup=0.3)
    o = orc.Forest([text])
    _, trees = pyref.parse_newicks([text])
    assert orc.lower_to_rows(o.naive(), o.num_trees) == pyref.naive_dist_matrix(trees)
    assert any(pyref.num_leaves(t) != len(pyref._leafset(t)) for t in trees)

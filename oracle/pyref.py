"""pyref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A second, independent, deliberately naive pure-Python restatement of the same reference functions as
phybin_oracle.cpp (RFDistance.hs:116-134, 168-248, 284-341, 352-361; PreProcessor.hs:21-32;
Parser.hs:57-110, 127-202).  It exists so that the C++ oracle is not its own only witness: the
golden matrices under tests/golden/ are produced by THIS file (tests/golden/make_golden.py) and
the C++ oracle must reproduce them, together with the reference's own known answer
(TestAll.hs:22-32).  Small inputs only (pure-Python loops).
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple

# A tree is ("L", label) or ("I", [children]).  IntSets are sorted tuples: tuple comparison is
# lexicographic on the ascending element list == Ord IntSet.


def _strip_ws(s: str) -> str:
    return re.sub(r"\s+", "", s)


class _P:
    RES = set("()[]:;,")

    def __init__(self, s: str, name_prefix: int):
        self.s, self.p, self.np = s, 0, name_prefix

    def at(self, c):
        return self.p < len(self.s) and self.s[self.p] == c

    def name(self):
        b = self.p
        while self.p < len(self.s) and self.s[self.p] not in self.RES:
            self.p += 1
        return self.s[b:self.p]

    def meta(self):  # Parser.hs:166-195
        if not self.at(":"):
            return
        self.p += 1
        m = re.compile(r"\d+(\.\d+)?e-?\d+").match(self.s, self.p) or \
            re.compile(r"-?\d+(\.\d+)?").match(self.s, self.p)
        if not m:
            raise ValueError("bad number at %d" % self.p)
        self.p = m.end()
        m = re.compile(r"\[\d+\]").match(self.s, self.p)
        if m:
            self.p = m.end()

    def subtree(self):
        if self.at("("):
            self.p += 1
            kids = []
            while True:
                k = self.subtree()
                self.meta()
                kids.append(k)
                if self.at(","):
                    self.p += 1
                    continue
                break
            if not self.at(")"):
                raise ValueError("expected ) at %d" % self.p)
            self.p += 1
            self.name()
            return ["I", kids]
        nm = self.name()
        if self.np >= 0:
            nm = nm[: self.np]
        return ["L", nm]

    def newick(self):
        t = self.subtree()
        self.meta()
        if not self.at(";"):
            raise ValueError("expected ; at %d" % self.p)
        self.p += 1
        return t


def parse_newicks(texts: Sequence[str], name_prefix: int = -1):
    """parseNewicks (Parser.hs:57-75): last file labelled first; children right-to-left."""
    seen, names = {}, []

    def assign(t):
        if t[0] == "L":
            nm = t[1]
            if nm not in seen:
                seen[nm] = len(names)
                names.append(nm)
            return ("L", seen[nm])
        kids = [None] * len(t[1])
        for i in range(len(t[1]) - 1, -1, -1):
            kids[i] = assign(t[1][i])
        return ("I", kids)

    per_file = [None] * len(texts)
    for fi in range(len(texts) - 1, -1, -1):
        s = _strip_ws(texts[fi])
        ps = _P(s, name_prefix)
        raw = [ps.newick()]
        while ps.p < len(s):
            raw.append(ps.newick())
        per_file[fi] = [assign(t) for t in raw]
    return names, [t for f in per_file for t in f]


def num_leaves(t) -> int:
    return 1 if t[0] == "L" else sum(num_leaves(k) for k in t[1])


def invert_dense(size: int, bip: Tuple[int, ...]) -> Tuple[int, ...]:
    s = set(bip)
    return tuple(ix for ix in range(size) if ix not in s)


def norm_bip(tot: int, bip: Tuple[int, ...]) -> Tuple[int, ...]:
    half = tot // 2
    flipped = invert_dense(tot, bip)
    if len(bip) < half:
        return bip
    if len(bip) > half:
        return flipped
    return min(bip, flipped)


def _leafset(t) -> Tuple[int, ...]:
    if t[0] == "L":
        return (t[1],)
    acc = set()
    for k in t[1]:
        acc.update(_leafset(k))
    return tuple(sorted(acc))


def all_bips(t) -> frozenset:
    size = num_leaves(t)
    out = set()

    def rec(n):
        if n[0] == "L":
            out.add(norm_bip(size, (n[1],)))
            return
        for k in n[1]:
            out.add(norm_bip(size, _leafset(k)))
        for k in n[1]:
            rec(k)

    rec(t)
    return frozenset(b for b in out if len(b) > 1)


def prune_tree_leaves(keep: set, t) -> Optional[tuple]:
    if t[0] == "L":
        return t if t[1] in keep else None
    kids = [x for x in (prune_tree_leaves(keep, k) for k in t[1]) if x is not None]
    if not kids:
        return None
    if len(kids) == 1:
        return kids[0]
    return ("I", kids)


def naive_dist_matrix(trees) -> List[List[int]]:
    labs = [set(_leafset(t)) for t in trees]
    each = [all_bips(t) for t in trees]
    mat = []
    for i in range(len(trees)):
        row = []
        for j in range(i):
            both = labs[i] & labs[j]
            if not both:
                row.append(0)
                continue
            bi = each[i] if len(both) == len(labs[i]) else all_bips(prune_tree_leaves(both, trees[i]))
            bj = each[j] if len(both) == len(labs[j]) else all_bips(prune_tree_leaves(both, trees[j]))
            row.append(len(bi - bj) + len(bj - bi))
        mat.append(row)
    return mat


def hash_rf(trees) -> List[List[int]]:
    T = len(trees)
    table = {}
    for ix, t in enumerate(trees):
        for b in all_bips(t):
            table.setdefault(b, set()).add(ix)
    mat = [[0] * i for i in range(T)]
    for have in table.values():
        for i in have:
            for j in range(T):
                if j in have:
                    continue
                if j < i:
                    mat[i][j] += 1
                else:
                    mat[j][i] += 1
    return mat


def print_dist_mat(mat: List[List[int]]) -> str:
    rule = "-" * 41 + "\n"
    o = "Robinson-Foulds distance (matrix format):\n" + rule
    for row in mat:
        o += "".join("%d " % e for e in row) + "0\n"
    return o + rule


# ---- strict consensus ----------------------------------------------------------------

def consensus_bips(trees) -> frozenset:
    """The bipartitions every tree has: foldl' S.intersection (allBips hd) (map allBips tl)
    (consensusTree, RFDistance.hs:396-400)."""
    if not trees:
        raise ValueError("Cannot take the consensusTree of the empty list")  # :397
    acc = all_bips(trees[0])
    for t in trees[1:]:
        acc = acc & all_bips(t)
    return acc


def bips_to_tree(num_taxa: int, bips) -> tuple:
    """bipsToTree (RFDistance.hs:410-445): start from the num_taxa singleton leaves; visit the bipartitions in
    increasing size (stable over Set order); the current subtrees that lie inside the bipartition are glommed,
    in list order, under a new interior node that goes to the FRONT of the list."""
    ordered = sorted(sorted(bips), key=len)  # L.sortBy (compare `on` bipSize) (S.toList origbip), stable
    subtrees = [(frozenset([ix]), ("L", ix)) for ix in range(num_taxa)]
    for bip in ordered:
        inside = set(bip)
        in_ = [st for st in subtrees if st[0] <= inside]   # isMatch bip x = denseIsSubset x bip
        out = [st for st in subtrees if not st[0] <= inside]
        if not in_:
            raise ValueError("bipsToTree: Internal error!  No match for bip: %r" % (bip,))
        union = frozenset().union(*[s for s, _ in in_])
        subtrees = [(union, ("I", [t for _, t in in_]))] + out
    if not subtrees:
        raise ValueError("bipsToTree: internal error")
    if len(subtrees) == 1:
        return subtrees[0][1]
    return ("I", [t for _, t in subtrees])


def consensus_tree(num_taxa: int, trees) -> tuple:
    return bips_to_tree(num_taxa, consensus_bips(trees))


def to_newick(names: Sequence[str], t) -> str:
    """displayStrippedTree (CoreTypes.hs:127-132) on one line: names and parentheses only, ", " separators
    (the reference lays the same tokens out with a pretty-printer; its parser drops all whitespace, Parser.hs:35)."""
    def rec(n):
        if n[0] == "L":
            return names[n[1]]
        return "(" + ", ".join(rec(k) for k in n[1]) + ")"
    return rec(t) + ";"


# ---- clustering consumer --------------------------------------------------------------

def flat_clusters(dist, n: int, linkage: str, thresh: float) -> List[int]:
    """sliceDendro thresh (C.dendrogram linkage items dist)  (PhyBin.hs:358,415-422) as labels[i] = smallest member
    of i's cluster.  C.dendrogram is hierarchical-clustering-0.4.6 (stack.yaml:9), NOT under /root/reference:
    PARITY UNPINNED.  Restated from the published package's distance-matrix algorithm, the naive way: find the
    minimal distance over all active pairs (k1 < k2), the first in lexicographic order wins ties; merge, the new
    cluster keeps the lower key; Lance-Williams update (single: min, complete: max, UPGMA: size-weighted mean).
    For these monotone linkages the flat clusters are what has merged before the first distance > thresh."""
    d = {(i, j): float(dist(i, j)) for i in range(n) for j in range(i + 1, n)}
    size = {i: 1 for i in range(n)}
    label = list(range(n))
    while len(size) > 1:
        keys = sorted(size)
        best, pair = None, None
        for x, i in enumerate(keys):
            for j in keys[x + 1:]:
                if best is None or d[(i, j)] < best:
                    best, pair = d[(i, j)], (i, j)
        if best > thresh:
            break
        a, b = pair
        for x in keys:
            if x in (a, b):
                continue
            da, db = d[(min(a, x), max(a, x))], d[(min(b, x), max(b, x))]
            if linkage == "single":
                v = min(da, db)
            elif linkage == "complete":
                v = max(da, db)
            else:
                v = (size[a] * da + size[b] * db) / (size[a] + size[b])
            d[(min(a, x), max(a, x))] = v
        size[a] += size.pop(b)
        label = [a if l == b else l for l in label]
    return label

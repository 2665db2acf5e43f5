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

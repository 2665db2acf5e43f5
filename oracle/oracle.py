"""ctypes loader for the C++ CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package phybin_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libphybin_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "phybin_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_parse.restype = C.c_void_p
        L.orc_parse.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_error.restype = C.c_char_p
        L.orc_error.argtypes = [C.c_void_p]
        for f in ("orc_num_trees", "orc_num_labels"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_label_name.restype = C.c_char_p
        L.orc_label_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_tree_num_leaves.argtypes = [C.c_void_p, C.c_int]
        L.orc_filter_num_leaves.argtypes = [C.c_void_p, C.c_int]
        L.orc_all_bips.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_all_bips_total_labels.argtypes = [C.c_void_p, C.c_int]
        L.orc_hashrf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_naive.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_print_dist_mat.restype = C.c_void_p
        L.orc_print_dist_mat.argtypes = [C.c_void_p, C.c_int]
        L.orc_free_str.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def max_threads():
    """Threads the "literal_all_cores" variant of Forest.hashrf uses."""
    f = lib().orc_max_threads
    f.restype = C.c_int
    return int(f())


class Forest:
    """Parsed trees + global label table, as parseNewicks returns them (Parser.hs:57)."""

    def __init__(self, texts: Sequence[str], name_prefix: int = -1):
        L = lib()
        arr = (C.c_char_p * len(texts))(*[t.encode() for t in texts])
        self._h = L.orc_parse(len(texts), arr, name_prefix)
        err = L.orc_error(self._h)
        if err:
            msg = err.decode()
            L.orc_free(self._h)
            self._h = None
            raise ValueError(msg)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_free(self._h)
            self._h = None

    @property
    def num_trees(self) -> int:
        return lib().orc_num_trees(self._h)

    @property
    def labels(self) -> List[str]:
        L = lib()
        return [L.orc_label_name(self._h, i).decode() for i in range(L.orc_num_labels(self._h))]

    def num_leaves(self, t: int) -> int:
        return lib().orc_tree_num_leaves(self._h, t)

    def filter_num_leaves(self, expected: int) -> int:
        return lib().orc_filter_num_leaves(self._h, expected)

    def all_bips(self, t: int) -> List[Tuple[int, ...]]:
        """allBips of tree t as ascending label tuples, in Set order (RFDistance.hs:247)."""
        L = lib()
        tot = L.orc_all_bips_total_labels(self._h, t)
        k = L.orc_all_bips(self._h, t, None, None, 0)
        sizes = np.zeros(max(k, 1), dtype=np.int32)
        labs = np.zeros(max(tot, 1), dtype=np.int32)
        k2 = L.orc_all_bips(self._h, t, sizes.ctypes.data, labs.ctypes.data, tot)
        assert k2 == k
        out, pos = [], 0
        for i in range(k):
            out.append(tuple(int(x) for x in labs[pos:pos + sizes[i]]))
            pos += sizes[i]
        return out

    def hashrf(self, num_taxa: int = 0, variant: str = "literal", stats: Optional[dict] = None) -> np.ndarray:
        """Flat lower-triangular matrix: (i, j), j < i at i(i-1)/2 + j (RFDistance.hs:284)."""
        T = self.num_trees
        out = np.zeros(T * (T - 1) // 2 if T else 0, dtype=np.int64)
        nu = C.c_int64(0)
        b, g = C.c_double(0), C.c_double(0)
        code = {"literal": 0, "closed": 1, "literal_all_cores": 2}[variant]
        rc = lib().orc_hashrf(self._h, num_taxa, code, out.ctypes.data, C.byref(nu), C.byref(b), C.byref(g))
        if rc:
            raise RuntimeError("oracle hashRF failed")
        if stats is not None:
            stats.update(n_unique=nu.value, build_s=b.value, ingest_s=g.value)
            stats.update(threads=max_threads() if code == 2 else 1)
        return out

    def consensus_bips(self, members: Sequence[int]) -> List[Tuple[int, ...]]:
        """The bipartitions of consensusTree over the given trees, in Set order:
        foldl' S.intersection (allBips hd) (map allBips tl)  (RFDistance.hs:396-400)."""
        members = list(members)
        if not members:
            raise ValueError("Cannot take the consensusTree of the empty list")  # :397
        acc = set(self.all_bips(members[0]))
        for t in members[1:]:
            acc &= set(self.all_bips(t))
        return sorted(acc)

    def naive(self) -> np.ndarray:
        """fst . naiveDistMatrix (RFDistance.hs:168), same flat layout as hashrf()."""
        T = self.num_trees
        out = np.zeros(T * (T - 1) // 2 if T else 0, dtype=np.int64)
        if lib().orc_naive(self._h, out.ctypes.data):
            raise RuntimeError("oracle naiveDistMatrix failed")
        return out


def lower_to_rows(flat: np.ndarray, T: int) -> List[List[int]]:
    """Flat lower triangle -> the reference's DistanceMatrix shape [[],[d10],[d20,d21],...]."""
    rows, pos = [], 0
    for i in range(T):
        rows.append([int(x) for x in flat[pos:pos + i]])
        pos += i
    return rows


def lower_to_packed_upper(flat: np.ndarray, T: int) -> np.ndarray:
    """Reference layout (row i, col j<i) -> packed row-major upper triangle used by the C ABI:
    pair (a, b), a < b at a*T - a(a+1)/2 + (b - a - 1)."""
    out = np.zeros_like(flat)
    if T < 2:
        return out
    ii, jj = np.tril_indices(T, -1)  # ii > jj, row-major over the lower triangle == flat order
    a, b = jj.astype(np.int64), ii.astype(np.int64)
    out[a * T - a * (a + 1) // 2 + (b - a - 1)] = flat
    return out


def print_dist_mat(flat: np.ndarray, T: int) -> str:
    L = lib()
    flat = np.ascontiguousarray(flat, dtype=np.int64)
    p = L.orc_print_dist_mat(flat.ctypes.data, T)
    s = C.string_at(p).decode()
    L.orc_free_str(p)
    return s

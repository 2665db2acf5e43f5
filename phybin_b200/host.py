"""Host side of the path: trees, the global label table and bipartition extraction.

Thin Python face of libphybin_host.so (C++).  Mirrors what PhyBin's driver does before it calls
hashRF / naiveDistMatrix (PhyBin.hs:125-194): parseNewicks, the numLeaves filter, allBips.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._native import host_lib, u32p, u64p

RANDOM_TOPOLOGY = 0
NNI_FAMILY = 1


def _take(ptr, n_words: int) -> np.ndarray:
    """Copies a malloc'd uint64 buffer into a numpy array and frees it."""
    L = host_lib()
    n = max(int(n_words), 0)
    arr = np.empty(n, dtype=np.uint64)
    if n:
        C.memmove(arr.ctypes.data, ptr, n * 8)
    L.phb_free_buf(ptr)
    return arr


class Forest:
    """A list of NewickTrees sharing one LabelTable (the `[FullTree]` parseNewicks returns)."""

    def __init__(self, handle):
        self._h = handle
        err = host_lib().phb_error(handle)
        if err:
            msg = err.decode()
            host_lib().phb_free(handle)
            self._h = None
            raise ValueError(msg)

    def __del__(self):
        if getattr(self, "_h", None):
            host_lib().phb_free(self._h)
            self._h = None

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def parse(cls, texts: Sequence[str], name_prefix: int = -1) -> "Forest":
        arr = (C.c_char_p * len(texts))(*[t.encode() for t in texts])
        return cls(host_lib().phb_parse(len(texts), arr, name_prefix))

    @classmethod
    def synthetic(cls, T: int, n: int, seed: int, family: int = RANDOM_TOPOLOGY, nni_max: int = 8,
                  p_missing: float = 0.0) -> "Forest":
        return cls(host_lib().phb_synth(T, n, seed & (2**64 - 1), family, nni_max, p_missing))

    # -- queries ------------------------------------------------------------------------------
    def __len__(self) -> int:
        return host_lib().phb_num_trees(self._h)

    @property
    def num_labels(self) -> int:
        return host_lib().phb_num_labels(self._h)

    @property
    def labels(self) -> List[str]:
        L = host_lib()
        return [L.phb_label_name(self._h, i).decode() for i in range(self.num_labels)]

    def num_leaves(self, t: int) -> int:
        return host_lib().phb_tree_num_leaves(self._h, t)

    def filter_num_leaves(self, expected: int) -> int:
        return host_lib().phb_filter_num_leaves(self._h, expected)

    def slice(self, t0: int, t1: int) -> int:
        return host_lib().phb_slice(self._h, t0, t1)

    def has_duplicate_leaves(self) -> bool:
        return bool(host_lib().phb_has_duplicate_leaves(self._h))

    def to_newick(self, t0: int = 0, t1: Optional[int] = None) -> str:
        L = host_lib()
        p = L.phb_to_newick(self._h, t0, len(self) if t1 is None else t1)
        s = C.string_at(p).decode()
        L.phb_free_buf(p)
        return s

    @property
    def words(self) -> int:
        return host_lib().phb_words(self._h)

    # -- device inputs ------------------------------------------------------------------------
    def tree_arrays(self):
        """(node_label int32[N], node_parent uint32[N], tree_begin uint64[T+1], num_leaves uint32[T]): the trees as
        the flat pre-order arrays prf_allbips takes (phb_tree_arrays).  Views into the forest: keep it alive and
        unmodified while they are in use."""
        L = host_lib()
        lp, pp, bp, nl = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_uint64(0)
        if L.phb_tree_arrays(self._h, C.byref(lp), C.byref(pp), C.byref(bp), C.byref(nl), C.byref(n)):
            raise RuntimeError("phb_tree_arrays failed")
        T, N = len(self), n.value

        def view(ptr, count, ctype, dtype):
            if not count:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,))

        return (view(lp, N, C.c_int32, np.int32), view(pp, N, C.c_uint32, np.uint32),
                view(bp, T + 1, C.c_uint64, np.uint64), view(nl, T, C.c_uint32, np.uint32))

    def all_bips(self, W: Optional[int] = None, threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """(bips[nnz, W] uint64, tree_off[T+1] uint64): allBips of every tree, bit-packed."""
        W = W or self.words
        bp, op, nnz = u64p(), u64p(), C.c_uint64(0)
        if host_lib().phb_allbips(self._h, W, threads, C.byref(bp), C.byref(op), C.byref(nnz)):
            raise RuntimeError("phb_allbips failed (W too small or out of memory)")
        return _take(bp, nnz.value * W).reshape(-1, W), _take(op, len(self) + 1)

    def raw_clades(self, W: Optional[int] = None, threads: int = 0):
        """(clades[nnz, W], tree_off[T+1], leafsets[T, W]) for the tolerant path."""
        W = W or self.words
        cp, op, lp, nnz = u64p(), u64p(), u64p(), C.c_uint64(0)
        if host_lib().phb_raw_clades(self._h, W, threads, C.byref(cp), C.byref(op), C.byref(lp),
                                     C.byref(nnz)):
            raise RuntimeError("phb_raw_clades failed (W too small or out of memory)")
        T = len(self)
        return (_take(cp, nnz.value * W).reshape(-1, W), _take(op, T + 1),
                _take(lp, T * W).reshape(T, W))


    def raw_clades_ex(self, W: Optional[int] = None):
        """phb_raw_clades_ex: the tolerant path's inputs for a forest whose trees may repeat a leaf label:
        (clades[nnz, W], outsets[nnz, W], tree_off[T+1], leafsets[T, W], dup_off[T+1], dup_label, dup_extra)."""
        L = host_lib()
        W = W or self.words
        T = len(self)
        c, o, off, ls, doff = u64p(), u64p(), u64p(), u64p(), u64p()
        dl, de = u32p(), u32p()
        nnz = C.c_uint64(0)
        rc = L.phb_raw_clades_ex(self._h, W, C.byref(c), C.byref(o), C.byref(off), C.byref(ls), C.byref(doff), C.byref(dl),
                                 C.byref(de), C.byref(nnz))
        if rc:
            raise ValueError("phb_raw_clades_ex failed (W too small?)")
        n = nnz.value
        off_a = _take(off, T + 1)
        doff_a = _take(doff, T + 1)
        nd = int(doff_a[T])
        dl_a = np.ctypeslib.as_array(dl, shape=(max(nd, 1),))[:nd].copy()
        de_a = np.ctypeslib.as_array(de, shape=(max(nd, 1),))[:nd].copy()
        L.phb_free_buf(dl)
        L.phb_free_buf(de)
        return (_take(c, n * W).reshape(n, W), _take(o, n * W).reshape(n, W), off_a, _take(ls, T * W).reshape(T, W),
                doff_a, dl_a, de_a)

    def consensus_newick(self, keys: np.ndarray, num_taxa: Optional[int] = None) -> str:
        """bipsToTree num_taxa keys (RFDistance.hs:410-445) printed on one line with this forest's label names
        (phb_consensus_newick).  keys: [k, W] bit sets, e.g. one cluster's range of RFContext.consensus()."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        W = keys.shape[1] if keys.ndim == 2 else self.words
        L = host_lib()
        p = L.phb_consensus_newick(self._h, keys.ctypes.data if keys.size else None, len(keys), W,
                                   self.num_labels if num_taxa is None else num_taxa)
        if not p:
            raise ValueError("bipsToTree: a bipartition matches no subtree (or names an unknown label)")
        s = C.string_at(p).decode()
        L.phb_free_buf(p)
        return s


LINKAGE = {"single": 0, "complete": 1, "UPGMA": 2}


def cluster_component(sub: np.ndarray, m: int, linkage: str, thresh: float) -> np.ndarray:
    """Flat clusters of one component under `linkage` cut at `thresh` (phb_cluster_component): sub is the packed
    upper triangle of the component's m x m distances (uint16 from the device, or float64); returns labels[m] =
    smallest local index of each item's cluster.  The part of C.dendrogram + sliceDendro (PhyBin.hs:358,415-422)
    that remains after the device split the trees into components."""
    L = host_lib()
    labels = np.empty(m, dtype=np.uint32)
    if sub.dtype == np.float64:
        sub = np.ascontiguousarray(sub)
        rc = L.phb_cluster_component_f64(sub.ctypes.data, m, LINKAGE[linkage], float(thresh), labels.ctypes.data)
    else:
        sub = np.ascontiguousarray(sub, dtype=np.uint16)
        rc = L.phb_cluster_component(sub.ctypes.data, m, LINKAGE[linkage], float(thresh), labels.ctypes.data)
    if rc:
        raise ValueError("phb_cluster_component: %s" % ("component too large for a dense matrix" if rc == -2 else "bad arguments"))
    return labels


def agglomerate(sub: np.ndarray, m: int, linkage: str, thresh: float = float("inf")):
    """phb_agglomerate: the whole agglomeration of m items from the packed upper triangle `sub` (uint16), cut at
    `thresh` (inf = full dendrogram).  Returns (labels[m], merge_a, merge_b, merge_d): step k merged the clusters with
    smallest members merge_a[k] < merge_b[k] at distance merge_d[k] (C.dendrogram, PhyBin.hs:358)."""
    L = host_lib()
    sub = np.ascontiguousarray(sub, dtype=np.uint16)
    labels = np.empty(m, dtype=np.uint32)
    ma = np.empty(max(m - 1, 1), dtype=np.uint32)
    mb = np.empty(max(m - 1, 1), dtype=np.uint32)
    md = np.empty(max(m - 1, 1), dtype=np.float64)
    n = C.c_uint32(0)
    rc = L.phb_agglomerate(sub.ctypes.data, m, LINKAGE[linkage], float(thresh), labels.ctypes.data, ma.ctypes.data,
                           mb.ctypes.data, md.ctypes.data, C.byref(n))
    if rc:
        raise ValueError("phb_agglomerate: %s" % ("too large for a dense matrix" if rc == -2 else "bad arguments"))
    return labels, ma[:n.value], mb[:n.value], md[:n.value]


def dendrogram_text(names: Sequence[str], merge_a, merge_b, merge_d) -> str:
    """phb_dendrogram_text: `show (fmap treename dendro)` as the reference writes it to dendrogram.txt
    (PhyBin.hs:238-243) from a full merge list."""
    L = host_lib()
    m = len(names)
    arr = (C.c_char_p * m)(*[n.encode() for n in names])
    ma = np.ascontiguousarray(merge_a, dtype=np.uint32)
    mb = np.ascontiguousarray(merge_b, dtype=np.uint32)
    md = np.ascontiguousarray(merge_d, dtype=np.float64)
    p = L.phb_dendrogram_text(arr, m, ma.ctypes.data, mb.ctypes.data, md.ctypes.data, len(ma))
    if not p:
        raise ValueError("phb_dendrogram_text: the merge list is not a full dendrogram of %d items" % m)
    try:
        return C.string_at(p).decode()
    finally:
        L.phb_free_buf(p)


def slice_dendrogram(m: int, merge_a, merge_b, merge_d, thresh: float) -> np.ndarray:
    """sliceDendro thresh dendro (PhyBin.hs:415-422) on a FULL merge list: walking down from the root, a Branch whose
    distance exceeds `thresh` is split into its children, anything else becomes one flat cluster.  Returns labels[m] =
    smallest member of each item's cluster.  (For the monotone linkages used here this equals stopping the agglomeration
    at the first distance > thresh, which is what prf_agglomerate / phb_agglomerate do with a finite thresh.)"""
    n = len(merge_a)
    if n + 1 != m:
        raise ValueError("slice_dendrogram needs the full merge list of %d items" % m)
    cur = list(range(m))            # node standing for each cluster key; node k < m: leaf, m + k: merge k
    left, right = [0] * n, [0] * n
    for k in range(n):
        a, b = int(merge_a[k]), int(merge_b[k])
        left[k], right[k] = cur[a], cur[b]
        cur[a] = m + k
    labels = np.empty(m, dtype=np.uint32)
    stack = [(m + n - 1 if n else 0, False)]
    while stack:                     # (node, inside a flat cluster already?)
        node, flat = stack.pop()
        if node < m:
            members = [node]
        elif not flat and merge_d[node - m] > thresh:
            stack.append((left[node - m], False))
            stack.append((right[node - m], False))
            continue
        else:
            members, todo = [], [node]
            while todo:
                x = todo.pop()
                if x < m:
                    members.append(x)
                else:
                    todo += [left[x - m], right[x - m]]
        labels[members] = min(members)
    return labels


def bits_to_labels(row: np.ndarray) -> Tuple[int, ...]:
    """One W-word bit set -> ascending label tuple (for tests and debugging)."""
    out = []
    for w, word in enumerate(row.tolist()):
        while word:
            b = (word & -word).bit_length() - 1
            out.append(w * 64 + b)
            word &= word - 1
    return tuple(out)

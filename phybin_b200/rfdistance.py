"""Host-side mirror of Bio.Phylogeny.PhyBin.RFDistance's public face for the distance-matrix path
(same names, argument meaning and results), computed by libphybin_rf.so on a B200.

    hashRF          :: Int -> [NewickTree a] -> DistanceMatrix                 RFDistance.hs:284
    naiveDistMatrix :: [NewickTree DefDecor] -> (DistanceMatrix, Vector (Set DenseLabelSet))   :168
    printDistMat    :: Handle -> DistanceMatrix -> IO ()                       :352

There is no CPU implementation in this module: every matrix comes from the CUDA library, and the
calls raise if it is missing or no device is present.
"""
from __future__ import annotations

import ctypes as C
from typing import IO, List, Optional, Tuple

import numpy as np

from . import _native
from ._native import prf_agglo_info, prf_consensus_info, prf_extract_info, prf_stats, rf_lib, host_lib
from .host import LINKAGE, Forest, bits_to_labels, cluster_component, dendrogram_text

VARIANT_AUTO, VARIANT_SPARSE, VARIANT_DENSE = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
EXTRACT_BIPS, EXTRACT_CLADES = 0, 1


class PhyBinRFError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"phybin_rf error {code}: {msg}")
        self.code = code


def packed_index(T: int, a, b):
    """p(a, b) for a < b: position of d(a, b) in the packed row-major upper triangle."""
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    return a * T - a * (a + 1) // 2 + (b - a - 1)


class DistanceMatrix:
    """`V.Vector (U.Vector Int)`: row i holds d(i, 0..i-1) (RFDistance.hs:156, :306-311), stored as
    the uint16 packed upper triangle the device produces."""

    def __init__(self, upper: np.ndarray, T: int):
        self.upper = upper
        self.T = T

    def __len__(self) -> int:
        return self.T

    def row(self, i: int) -> np.ndarray:
        j = np.arange(i, dtype=np.int64)
        return self.upper[packed_index(self.T, j, i)].astype(np.int64)

    def __getitem__(self, i: int) -> np.ndarray:
        return self.row(i)

    def rows(self) -> List[List[int]]:
        return [self.row(i).tolist() for i in range(self.T)]

    def lower_flat(self) -> np.ndarray:
        """(i, j), j < i at i(i-1)/2 + j -- the reference's rows back to back."""
        T = self.T
        if T < 2:
            return np.zeros(0, dtype=np.int64)
        ii, jj = np.tril_indices(T, -1)
        return self.upper[packed_index(T, jj, ii)].astype(np.int64)

    def save_bin(self, path: str, chunk: int = 1 << 26) -> None:
        """Binary form (phb_dist_bin_*: 32-byte header + packed uint16 upper triangle), written in chunks."""
        L = host_lib()
        if L.phb_dist_bin_begin(str(path).encode(), self.T):
            raise OSError(f"cannot write {path}")
        up = np.ascontiguousarray(self.upper, dtype=np.uint16)
        for b in range(0, len(up), chunk):
            part = up[b:b + chunk]
            if L.phb_dist_bin_append(str(path).encode(), part.ctypes.data, len(part)):
                raise OSError(f"cannot write {path}")

    @classmethod
    def load_bin(cls, path: str) -> "DistanceMatrix":
        L = host_lib()
        T, p = C.c_uint32(0), C.c_void_p()
        if L.phb_dist_bin_read(str(path).encode(), C.byref(T), C.byref(p)):
            raise OSError(f"{path} is not a phybin distance matrix file")
        n = T.value * (T.value - 1) // 2 if T.value else 0
        up = np.empty(n, dtype=np.uint16)
        if n:
            C.memmove(up.ctypes.data, p, n * 2)
        L.phb_free_buf(p)
        return cls(up, T.value)

    def __repr__(self) -> str:  # `show` of the Haskell value for small matrices
        return str(self.rows()).replace(" ", "")


class MultiRF:
    """prf_multi: one process driving several GPUs of the box (the trees are dealt to the GPUs, the build is
    hash-partitioned over them, every GPU computes and returns a slice of the matrix)."""

    def __init__(self, n_gpus: int = 0, devices=None):
        self._L = rf_lib()
        arr = (C.c_int * len(devices))(*devices) if devices else None
        self._m = self._L.prf_multi_create(len(devices) if devices else n_gpus, arr)
        if not self._m:
            raise PhyBinRFError(-2, "prf_multi_create failed (no CUDA device, or more than 16 GPUs)")

    @property
    def n_gpus(self) -> int:
        return int(self._L.prf_multi_gpus(self._m))

    def close(self):
        if getattr(self, "_m", None):
            self._L.prf_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        self.close()

    def hashrf(self, bips: np.ndarray, tree_off: np.ndarray, editdist: int = -1):
        """prf_multi_hashrf on host arrays: (upper uint16[P], mask uint32[...] | None, stats)."""
        bips = np.ascontiguousarray(bips, dtype=np.uint64)
        tree_off = np.ascontiguousarray(tree_off, dtype=np.uint64)
        T = len(tree_off) - 1
        W = bips.shape[1] if bips.ndim == 2 else 1
        P = T * (T - 1) // 2 if T else 0
        out = np.empty(P, dtype=np.uint16)
        mask = np.zeros((P + 31) // 32, dtype=np.uint32) if editdist >= 0 else None
        st = prf_stats()
        rc = self._L.prf_multi_hashrf(self._m, bips.ctypes.data, tree_off.ctypes.data, T, W, editdist, out.ctypes.data,
                                      mask.ctypes.data if mask is not None else None, C.byref(st))
        if rc != 0:
            raise PhyBinRFError(rc, self._L.prf_multi_last_error(self._m).decode())
        self.T = T
        return out, mask, st.as_dict()

    def components(self) -> Tuple[np.ndarray, int]:
        """prf_multi_components after hashrf(editdist >= 0): every GPU's union-find over its resident mask shard,
        merged on GPU 0 -- (labels[T] = smallest tree id of each tree's component, number of components)."""
        labels = np.empty(self.T, dtype=np.uint32)
        n = C.c_uint32(0)
        rc = self._L.prf_multi_components(self._m, labels.ctypes.data, C.byref(n))
        if rc != 0:
            raise PhyBinRFError(rc, self._L.prf_multi_last_error(self._m).decode())
        return labels, n.value


class RFContext:
    """One prf_ctx: a CUDA device, a stream, and the device-resident state of one problem."""

    def __init__(self, device: int = 0):
        self._L = rf_lib()
        self._h = self._L.prf_create(device)
        if not self._h:
            raise PhyBinRFError(-2, self._L.prf_last_error(None).decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.prf_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc: int):
        if rc != 0:
            raise PhyBinRFError(rc, self._L.prf_last_error(self._h).decode())

    def debug_small_tiles(self, on: bool = True):
        """Test hook: run the sparse matrix kernel with 256-entry tiles (prf_debug_small_tiles)."""
        self._ck(self._L.prf_debug_small_tiles(self._h, 1 if on else 0))

    @property
    def launches(self) -> int:
        return int(self._L.prf_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(self._L.prf_stream(self._h) or 0)

    # ---- hashRF ------------------------------------------------------------------------------
    def hashrf(self, bips: np.ndarray, tree_off: np.ndarray, editdist: int = -1,
               fetch: bool = True) -> Tuple[Optional[np.ndarray], Optional[np.ndarray], dict]:
        """One-shot prf_hashrf on host arrays.  Returns (upper uint16[P], mask uint32[...] | None, stats)."""
        bips = np.ascontiguousarray(bips, dtype=np.uint64)
        tree_off = np.ascontiguousarray(tree_off, dtype=np.uint64)
        T = len(tree_off) - 1
        W = bips.shape[1] if bips.ndim == 2 else 1
        P = T * (T - 1) // 2 if T else 0
        out = np.empty(P, dtype=np.uint16) if fetch else None
        mask = np.empty((P + 31) // 32, dtype=np.uint32) if (fetch and editdist >= 0) else None
        st = prf_stats()
        self._ck(self._L.prf_hashrf(self._h, bips.ctypes.data, tree_off.ctypes.data, T, W, editdist,
                                    out.ctypes.data if out is not None else None,
                                    mask.ctypes.data if mask is not None else None, C.byref(st)))
        return out, mask, st.as_dict()

    def hashrf_load(self, bips, tree_off, T: Optional[int] = None, W: Optional[int] = None,
                    kind: int = MEM_HOST) -> dict:
        """prf_hashrf_load.  With kind=MEM_DEVICE pass raw device pointers (ints) plus T and W."""
        st = prf_stats()
        if kind == MEM_HOST:
            bips = np.ascontiguousarray(bips, dtype=np.uint64)
            tree_off = np.ascontiguousarray(tree_off, dtype=np.uint64)
            T = len(tree_off) - 1
            W = bips.shape[1] if bips.ndim == 2 else 1
            self._keep = (bips, tree_off)
            bp, op = bips.ctypes.data, tree_off.ctypes.data
        else:
            bp, op = int(bips), int(tree_off)
        self._ck(self._L.prf_hashrf_load(self._h, bp, op, T, W, kind, C.byref(st)))
        self.T = T
        self.P = T * (T - 1) // 2 if T else 0
        return st.as_dict()

    def hashrf_compute(self, p_begin: int = 0, p_end: Optional[int] = None, editdist: int = -1,
                       variant: int = VARIANT_AUTO, d_out: int = 0, d_mask: int = 0,
                       want_stats: bool = True) -> Optional[dict]:
        p_end = self.P if p_end is None else p_end
        st = prf_stats()
        self._ck(self._L.prf_hashrf_compute(self._h, p_begin, p_end, editdist, variant, d_out or None,
                                            d_mask or None, C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    # ---- hashRF over several GPUs (one context per GPU; phybin_b200.distributed has the torch.distributed driver) ----
    def team_init(self, rank: int, world: int, T: int, W: int, nnz_total: int, max_local_entries: int,
                  max_tree_entries: int) -> bytes:
        """prf_team_init: allocates this rank's exchange arena; returns its handle (pass all ranks' handles, in rank
        order, to team_connect)."""
        buf = C.create_string_buffer(self._L.prf_team_handle_bytes())
        self._ck(self._L.prf_team_init(self._h, rank, world, T, W, nnz_total, max_local_entries, max_tree_entries, buf))
        return buf.raw

    def team_connect(self, handles: bytes):
        self._ck(self._L.prf_team_connect(self._h, handles))

    def team_load(self, d_bips: int, d_tree_off: int, T_local: int, T: int) -> dict:
        """prf_team_load (collective): device pointers of THIS rank's trees' bipartitions and offsets."""
        st = prf_stats()
        self._ck(self._L.prf_team_load(self._h, d_bips, d_tree_off, T_local, C.byref(st)))
        self.T = T
        self.P = T * (T - 1) // 2 if T else 0
        return st.as_dict()

    # ---- tolerant ----------------------------------------------------------------------------
    def tolerant(self, clades: np.ndarray, tree_off: np.ndarray, leafsets: np.ndarray, editdist: int = -1,
                 fetch: bool = True):
        clades = np.ascontiguousarray(clades, dtype=np.uint64)
        tree_off = np.ascontiguousarray(tree_off, dtype=np.uint64)
        leafsets = np.ascontiguousarray(leafsets, dtype=np.uint64)
        T = len(tree_off) - 1
        W = leafsets.shape[1] if leafsets.ndim == 2 else 1
        P = T * (T - 1) // 2 if T else 0
        out = np.empty(P, dtype=np.uint16) if fetch else None
        mask = np.empty((P + 31) // 32, dtype=np.uint32) if (fetch and editdist >= 0) else None
        st = prf_stats()
        self._ck(self._L.prf_tolerant(self._h, clades.ctypes.data, tree_off.ctypes.data, leafsets.ctypes.data,
                                      T, W, editdist, out.ctypes.data if out is not None else None,
                                      mask.ctypes.data if mask is not None else None, C.byref(st)))
        return out, mask, st.as_dict()

    def tolerant_load(self, clades, tree_off, leafsets, T: Optional[int] = None, W: Optional[int] = None,
                      kind: int = MEM_HOST) -> dict:
        st = prf_stats()
        if kind == MEM_HOST:
            clades = np.ascontiguousarray(clades, dtype=np.uint64)
            tree_off = np.ascontiguousarray(tree_off, dtype=np.uint64)
            leafsets = np.ascontiguousarray(leafsets, dtype=np.uint64)
            T = len(tree_off) - 1
            W = leafsets.shape[1] if leafsets.ndim == 2 else 1
            cp, op, lp = clades.ctypes.data, tree_off.ctypes.data, leafsets.ctypes.data
        else:
            cp, op, lp = int(clades), int(tree_off), int(leafsets)
        self._ck(self._L.prf_tolerant_load(self._h, cp, op, lp, T, W, kind, C.byref(st)))
        self.T = T
        self.P = T * (T - 1) // 2 if T else 0
        return st.as_dict()

    def tolerant_load_dups(self, clades, outsets, tree_off, leafsets, dup_off, dup_label, dup_extra) -> dict:
        """prf_tolerant_load_dups (host arrays, as Forest.raw_clades_ex returns them): trees may repeat a leaf label."""
        st = prf_stats()
        clades = np.ascontiguousarray(clades, dtype=np.uint64)
        outsets = np.ascontiguousarray(outsets, dtype=np.uint64)
        tree_off = np.ascontiguousarray(tree_off, dtype=np.uint64)
        leafsets = np.ascontiguousarray(leafsets, dtype=np.uint64)
        dup_off = np.ascontiguousarray(dup_off, dtype=np.uint64)
        dup_label = np.ascontiguousarray(dup_label, dtype=np.uint32)
        dup_extra = np.ascontiguousarray(dup_extra, dtype=np.uint32)
        T = len(tree_off) - 1
        W = leafsets.shape[1] if leafsets.ndim == 2 else 1
        self._ck(self._L.prf_tolerant_load_dups(self._h, clades.ctypes.data, outsets.ctypes.data, tree_off.ctypes.data,
                                                leafsets.ctypes.data, dup_off.ctypes.data, dup_label.ctypes.data,
                                                dup_extra.ctypes.data, T, W, MEM_HOST, C.byref(st)))
        self.T = T
        self.P = T * (T - 1) // 2 if T else 0
        return st.as_dict()

    def tolerant_compute(self, p_begin: int = 0, p_end: Optional[int] = None, editdist: int = -1,
                         d_out: int = 0, d_mask: int = 0, want_stats: bool = True) -> Optional[dict]:
        p_end = self.P if p_end is None else p_end
        st = prf_stats()
        self._ck(self._L.prf_tolerant_compute(self._h, p_begin, p_end, editdist, d_out or None, d_mask or None,
                                              C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    # ---- bipartition extraction on the device ------------------------------------------------------
    def allbips(self, trees: Forest, mode: int = EXTRACT_BIPS, W: Optional[int] = None) -> dict:
        """prf_allbips on a Forest's flat arrays: allBips (mode EXTRACT_BIPS) or the tolerant path's raw clades
        (EXTRACT_CLADES) of every tree, left on the device.  Returns the prf_extract_info fields."""
        label, parent, begin, leaves = trees.tree_arrays()
        W = W or trees.words
        info = prf_extract_info()
        self._ck(self._L.prf_allbips(self._h, label.ctypes.data if label.size else None,
                                     parent.ctypes.data if parent.size else None, begin.ctypes.data,
                                     leaves.ctypes.data if leaves.size else None, len(trees), W, MEM_HOST, mode,
                                     C.byref(info)))
        self._ext = (len(trees), W, mode, info.nnz)
        return info.as_dict()

    def allbips_result(self) -> Tuple[int, int, int, int]:
        """(d_keys, d_tree_off, d_leafsets, nnz): device pointers of the last extraction (prf_allbips_result)."""
        k, o, l = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_uint64(0)
        self._ck(self._L.prf_allbips_result(self._h, C.byref(k), C.byref(o), C.byref(l), C.byref(n)))
        return k.value or 0, o.value or 0, l.value or 0, n.value

    def allbips_fetch(self):
        """Host copies of the last extraction: (keys[nnz, W], tree_off[T+1]) or, in mode EXTRACT_CLADES,
        (clades[nnz, W], tree_off[T+1], leafsets[T, W])."""
        if getattr(self, "_ext", None) is None:
            raise PhyBinRFError(-5, "no extraction result on this context (call allbips first)")
        T, W, mode, nnz = self._ext
        keys = np.empty((nnz, W), dtype=np.uint64)
        off = np.empty(T + 1, dtype=np.uint64)
        leaf = np.empty((T, W), dtype=np.uint64) if mode == EXTRACT_CLADES else None
        self._ck(self._L.prf_allbips_fetch(self._h, keys.ctypes.data, off.ctypes.data,
                                           leaf.ctypes.data if leaf is not None else None))
        return (keys, off) if leaf is None else (keys, off, leaf)

    def hashrf_trees(self, trees: Forest, editdist: int = -1, fetch: bool = True, tolerant: bool = False,
                     pinned: Optional[tuple] = None):
        """One-shot prf_hashrf_trees / prf_tolerant_trees: parsed trees in, matrix out (allBips runs on the device).
        Returns (upper | None, mask | None, stats) like hashrf().  `pinned`: optional (upper, mask) host arrays to fill."""
        label, parent, begin, leaves = trees.tree_arrays()
        T, W = len(trees), trees.words
        P = T * (T - 1) // 2 if T else 0
        out = pinned[0] if pinned else (np.empty(P, dtype=np.uint16) if fetch else None)
        mask = pinned[1] if pinned else (np.empty((P + 31) // 32, dtype=np.uint32) if (fetch and editdist >= 0) else None)
        st, info = prf_stats(), prf_extract_info()
        fn = self._L.prf_tolerant_trees if tolerant else self._L.prf_hashrf_trees
        self._ck(fn(self._h, label.ctypes.data if label.size else None, parent.ctypes.data if parent.size else None,
                    begin.ctypes.data, leaves.ctypes.data if leaves.size else None, T, W, editdist,
                    out.ctypes.data if out is not None else None, mask.ctypes.data if mask is not None else None,
                    C.byref(st), C.byref(info)))
        self.T, self.P = T, P
        self._ext = (T, W, EXTRACT_CLADES if tolerant else EXTRACT_BIPS, info.nnz)
        d = st.as_dict()
        d["extract"] = info.as_dict()
        return out, mask, d

    def hashrf_load_trees(self, trees: Forest, W: Optional[int] = None) -> dict:
        """allBips on the device, then the hashRF build on its device-resident output (no host bipartitions)."""
        info = self.allbips(trees, EXTRACT_BIPS, W)
        k, o, _, _ = self.allbips_result()
        st = self.hashrf_load(k, o, T=len(trees), W=self._ext[1], kind=MEM_DEVICE)
        st["extract"] = info
        return st

    def tolerant_load_trees(self, trees: Forest, W: Optional[int] = None) -> dict:
        info = self.allbips(trees, EXTRACT_CLADES, W)
        k, o, l, _ = self.allbips_result()
        st = self.tolerant_load(k, o, l, T=len(trees), W=self._ext[1], kind=MEM_DEVICE)
        st["extract"] = info
        return st

    # ---- flat clusters --------------------------------------------------------------------------
    def mask_components(self, d_mask: int = 0) -> Tuple[np.ndarray, int]:
        """prf_mask_components on the context-owned mask of the last compute (or a device mask pointer):
        (labels[T] = smallest tree id of each tree's component of {d <= editdist}, number of components)."""
        labels = np.empty(self.T, dtype=np.uint32)
        n = C.c_uint32(0)
        self._ck(self._L.prf_mask_components(self._h, d_mask or None, self.T, labels.ctypes.data, C.byref(n)))
        return labels, n.value

    def mask_components_shard(self, p_begin: int, p_end: int, seeds=None, d_mask: int = 0) -> Tuple[np.ndarray, int]:
        """prf_mask_components_shard: union-find over the mask bits of the packed range [p_begin, p_end) (the
        context-owned mask of a sharded compute, or a device pointer), merged with the label arrays `seeds`
        ([n, T]) that other shards returned; an empty range merges the seeds only."""
        labels = np.empty(self.T, dtype=np.uint32)
        n = C.c_uint32(0)
        seeds = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32).reshape(-1, self.T)
        self._ck(self._L.prf_mask_components_shard(self._h, d_mask or None, p_begin, p_end, self.T,
                                                   seeds.ctypes.data if seeds is not None else None,
                                                   len(seeds) if seeds is not None else 0, labels.ctypes.data, C.byref(n)))
        return labels, n.value

    def fetch_submatrix(self, ids, d_matrix: int = 0) -> np.ndarray:
        """prf_fetch_submatrix: packed upper triangle of d(ids[i], ids[j]) for ascending tree ids, read from the
        context-owned result or from a device buffer holding the whole packed triangle (d_matrix)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        m = len(ids)
        out = np.empty(m * (m - 1) // 2, dtype=np.uint16)
        self._ck(self._L.prf_fetch_submatrix(self._h, d_matrix or None, self.T, ids.ctypes.data, m, out.ctypes.data))
        return out

    def agglomerate(self, ids=None, linkage: str = "UPGMA", thresh: float = float("inf"), d_matrix: int = 0):
        """prf_agglomerate: C.dendrogram linkage trees dist (PhyBin.hs:358) on the device for the ascending tree ids
        `ids` (None = every tree), cut at `thresh` (inf = the full dendrogram).  Returns (labels[m] = smallest local
        index of each tree's cluster, merge_a, merge_b, merge_d, info): step k merged the clusters with smallest local
        members merge_a[k] < merge_b[k] at distance merge_d[k]."""
        if ids is None:
            m, ids_p = self.T, None
        else:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            m, ids_p = len(ids), ids.ctypes.data
        labels = np.empty(max(m, 1), dtype=np.uint32)
        ma = np.empty(max(m - 1, 1), dtype=np.uint32)
        mb = np.empty(max(m - 1, 1), dtype=np.uint32)
        md = np.empty(max(m - 1, 1), dtype=np.float64)
        info = prf_agglo_info()
        self._ck(self._L.prf_agglomerate(self._h, d_matrix or None, self.T, ids_p, m, LINKAGE[linkage], float(thresh),
                                         labels.ctypes.data, ma.ctypes.data, mb.ctypes.data, md.ctypes.data, C.byref(info)))
        n = info.n_merges
        return labels[:m], ma[:n], mb[:n], md[:n], info.as_dict()

    def flat_clusters(self, editdist: int, linkage: str = "UPGMA", d_mask: int = 0, d_matrix: int = 0,
                      device_min: int = 64) -> Tuple[np.ndarray, int]:
        """sliceDendro editdist (C.dendrogram linkage trees dist)  (PhyBin.hs:358,415-422) as flat labels: labels[t] =
        smallest tree id of t's cluster.  Needs a full-triangle compute with this editdist (its mask; context-owned,
        or the caller's device buffers d_mask / d_matrix).  The device splits the trees into the connected components
        of {d <= editdist}; single linkage is done there.  For the other linkages every component of at least
        `device_min` trees is agglomerated on the device (prf_agglomerate), smaller ones on the host from their fetched
        sub-matrix (cluster_component: same rule, same arithmetic)."""
        labels, n = self.mask_components(d_mask)
        if linkage == "single":
            return labels, n
        order = np.argsort(labels, kind="stable")
        bounds = np.flatnonzero(np.diff(labels[order])) + 1
        out = labels.copy()
        for ids in np.split(order, bounds):
            if len(ids) < 3:  # a component of two trees is one cluster under every linkage
                continue
            ids = np.sort(ids).astype(np.uint32)
            if len(ids) >= device_min:
                local = self.agglomerate(ids, linkage, editdist, d_matrix)[0]
            else:
                local = cluster_component(self.fetch_submatrix(ids, d_matrix), len(ids), linkage, editdist)
            out[ids] = ids[local]
        return out, int(np.count_nonzero(out == np.arange(len(out))))

    def dendrogram(self, names, linkage: str = "UPGMA", d_matrix: int = 0) -> str:
        """dendrogram.txt (PhyBin.hs:238-243): the full agglomeration of every tree on the device, shown as the
        reference's `show (fmap treename dendro)`."""
        _, ma, mb, md, _ = self.agglomerate(None, linkage, float("inf"), d_matrix)
        return dendrogram_text(names, ma, mb, md)

    def consensus(self, labels: np.ndarray, bips=None, tree_off=None) -> Tuple[np.ndarray, np.ndarray, dict]:
        """prf_consensus: for every cluster (labels[t] = smallest tree id of t's cluster, as mask_components returns)
        the bipartitions all its trees share -- the intersection inside consensusTree (RFDistance.hs:396-400).
        Works on host arrays (bips[nnz, W], tree_off) or, by default, on the last device extraction
        (allbips / hashrf_load_trees).  Returns (keys[n_out, W], cons_off[T+1], info): cluster c's set is
        keys[cons_off[c]:cons_off[c+1]]; the ranges of trees that do not name a cluster are empty."""
        labels = np.ascontiguousarray(labels, dtype=np.uint32)
        T = len(labels)
        info = prf_consensus_info()
        if bips is None:
            if getattr(self, "_ext", None) is None:
                raise PhyBinRFError(-5, "no device extraction on this context: pass bips / tree_off, or call allbips first")
            eT, W, mode, _ = self._ext
            if mode != EXTRACT_BIPS or eT != T:
                raise PhyBinRFError(-5, "the last device extraction is not the allBips of these trees")
            k, o, _, _ = self.allbips_result()
            self._ck(self._L.prf_consensus(self._h, k, o, T, W, MEM_DEVICE, labels.ctypes.data, C.byref(info)))
        else:
            bips = np.ascontiguousarray(bips, dtype=np.uint64)
            tree_off = np.ascontiguousarray(tree_off, dtype=np.uint64)
            W = bips.shape[1] if bips.ndim == 2 else 1
            self._ck(self._L.prf_consensus(self._h, bips.ctypes.data, tree_off.ctypes.data, T, W, MEM_HOST,
                                           labels.ctypes.data, C.byref(info)))
        keys = np.empty((info.n_out, W), dtype=np.uint64)
        off = np.empty(T + 1, dtype=np.uint64)
        self._ck(self._L.prf_consensus_fetch(self._h, keys.ctypes.data, off.ctypes.data))
        return keys, off, info.as_dict()

    # ---- results -----------------------------------------------------------------------------
    def fetch(self, p_begin: int = 0, p_end: Optional[int] = None) -> np.ndarray:
        p_end = self.P if p_end is None else p_end
        out = np.empty(p_end - p_begin, dtype=np.uint16)
        self._ck(self._L.prf_fetch(self._h, p_begin, p_end, out.ctypes.data))
        return out

    def fetch_mask(self, p_begin: int = 0, p_end: Optional[int] = None) -> np.ndarray:
        p_end = self.P if p_end is None else p_end
        out = np.empty((p_end - p_begin + 31) // 32, dtype=np.uint32)
        self._ck(self._L.prf_fetch_mask(self._h, p_begin, p_end, out.ctypes.data))
        return out

    def fetch_rows(self, row0: int, row1: int) -> np.ndarray:
        T = self.T
        n = sum(T - 1 - a for a in range(row0, min(row1, T)))
        out = np.empty(n, dtype=np.uint16)
        self._ck(self._L.prf_fetch_rows(self._h, row0, row1, out.ctypes.data))
        return out


_default_ctx: Optional[RFContext] = None


def default_context() -> RFContext:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = RFContext(0)
    return _default_ctx


def shard_range(T: int, rank: int, world: int) -> Tuple[int, int]:
    """Aligned, balanced slice of the packed triangle for `rank` of `world` (prf_shard_range)."""
    b, e = C.c_uint64(0), C.c_uint64(0)
    rc = rf_lib().prf_shard_range(T, rank, world, C.byref(b), C.byref(e))
    if rc:
        raise PhyBinRFError(rc, "prf_shard_range: bad arguments")
    return b.value, e.value


# ---- the reference's names ---------------------------------------------------------------------

def hashRF(num_taxa: int, trees: Forest, ctx: Optional[RFContext] = None, extract: str = "device") -> DistanceMatrix:
    """hashRF num_taxa trees (RFDistance.hs:284).  `num_taxa` is accepted and ignored exactly as in
    the reference's IntSet build (it only reaches mkSingleDense, :295/:119).  The caller is
    expected to have applied the numLeaves filter (PhyBin.hs:187-194; Forest.filter_num_leaves)."""
    del num_taxa
    ctx = ctx or default_context()
    if extract == "device":  # allBips runs on the device too (prf_allbips)
        ctx.hashrf_load_trees(trees)
    elif extract == "host":  # the C++ host extraction (phb_allbips), e.g. for trees too large for prf_allbips
        ctx.hashrf_load(*trees.all_bips())
    else:
        raise ValueError("extract must be 'device' or 'host'")
    ctx.hashrf_compute(want_stats=False)
    return DistanceMatrix(ctx.fetch(), len(trees))


def naiveDistMatrix(trees: Forest, ctx: Optional[RFContext] = None):
    """naiveDistMatrix trees (RFDistance.hs:168): (matrix, per-tree allBips as sets of label tuples)."""
    ctx = ctx or default_context()
    ctx.allbips(trees, EXTRACT_BIPS)  # the second component of the result: every tree's allBips
    bips, boff = ctx.allbips_fetch()
    if trees.has_duplicate_leaves():
        # numLeaves counts a repeated label twice (CoreTypes.hs:289-291): per-tree sizes and outside sets from the host
        ctx.tolerant_load_dups(*trees.raw_clades_ex())
    else:
        ctx.tolerant_load_trees(trees)
    ctx.tolerant_compute(want_stats=False)
    upper = ctx.fetch()
    eachbips = [frozenset(bits_to_labels(b) for b in bips[int(boff[t]):int(boff[t + 1])])
                for t in range(len(trees))]
    return DistanceMatrix(upper, len(trees)), eachbips


def printDistMat(handle: IO[str], mat: DistanceMatrix) -> None:
    """printDistMat h mat (RFDistance.hs:352-361), byte-identical text."""
    L = host_lib()
    up = np.ascontiguousarray(mat.upper, dtype=np.uint16)
    p = L.phb_print_dist_mat(up.ctypes.data, mat.T)
    handle.write(C.string_at(p).decode())
    L.phb_free_buf(p)

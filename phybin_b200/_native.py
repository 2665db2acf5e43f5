"""ctypes bindings of the two in-tree native libraries.

libphybin_host.so (C++ host side) loads anywhere.  libphybin_rf.so (CUDA) is the only compute
path: if it is missing, fails to load, or finds no CUDA device, the callers raise -- there is no
CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u16p = C.POINTER(C.c_uint16)


class prf_extract_info(C.Structure):
    _fields_ = [("nnz", C.c_uint64), ("max_nodes", C.c_uint32), ("gpu_launches", C.c_int32),
                ("ms_h2d", C.c_float), ("ms_kernels", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class prf_consensus_info(C.Structure):
    _fields_ = [("n_out", C.c_uint64), ("n_clusters", C.c_uint32), ("gpu_launches", C.c_int32),
                ("ms_kernels", C.c_float), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class prf_agglo_info(C.Structure):
    _fields_ = [("n_merges", C.c_uint32), ("n_clusters", C.c_uint32), ("gpu_launches", C.c_int32),
                ("ms_kernels", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class prf_stats(C.Structure):
    _fields_ = [
        ("n_trees", C.c_uint64), ("n_words", C.c_uint64), ("nnz", C.c_uint64),
        ("n_unique", C.c_uint64), ("n_shared", C.c_uint64), ("n_shared_nnz", C.c_uint64),
        ("n_hits", C.c_uint64), ("n_pairs", C.c_uint64),
        ("max_bips", C.c_uint32), ("min_bips", C.c_uint32),
        ("variant_used", C.c_int32), ("gpu_launches", C.c_int32),
        ("ms_h2d", C.c_float), ("ms_dedup", C.c_float), ("ms_incidence", C.c_float),
        ("ms_matrix", C.c_float), ("ms_d2h", C.c_float), ("ms_total", C.c_float),
        ("n_flipped", C.c_uint32), ("reserved", C.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_host = None
_rf = None


def host_lib():
    global _host
    if _host is None:
        path = _build.HOST_SO
        if not os.path.exists(path):
            _build.build_host()
        L = C.CDLL(path)
        vp = C.c_void_p
        L.phb_parse.restype = vp
        L.phb_parse.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int]
        L.phb_synth.restype = vp
        L.phb_synth.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32, C.c_double]
        L.phb_free.argtypes = [vp]
        L.phb_error.restype = C.c_char_p
        L.phb_error.argtypes = [vp]
        for f in ("phb_num_trees", "phb_num_labels", "phb_words"):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [vp]
        L.phb_label_name.restype = C.c_char_p
        L.phb_label_name.argtypes = [vp, C.c_uint32]
        L.phb_tree_num_leaves.restype = C.c_uint32
        L.phb_tree_num_leaves.argtypes = [vp, C.c_uint32]
        L.phb_file_num_trees.restype = C.c_uint32
        L.phb_file_num_trees.argtypes = [vp, C.c_uint32]
        L.phb_filter_num_leaves.restype = C.c_uint32
        L.phb_filter_num_leaves.argtypes = [vp, C.c_uint32]
        L.phb_slice.restype = C.c_uint32
        L.phb_slice.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.phb_has_duplicate_leaves.argtypes = [vp]
        L.phb_to_newick.restype = vp
        L.phb_to_newick.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.phb_allbips.argtypes = [vp, C.c_uint32, C.c_int, C.POINTER(u64p), C.POINTER(u64p), u64p]
        L.phb_raw_clades.argtypes = [vp, C.c_uint32, C.c_int, C.POINTER(u64p), C.POINTER(u64p),
                                     C.POINTER(u64p), u64p]
        L.phb_raw_clades_ex.argtypes = [vp, C.c_uint32, C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u64p),
                                        C.POINTER(u64p), C.POINTER(u32p), C.POINTER(u32p), u64p]
        L.phb_tree_arrays.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), u64p]
        L.phb_consensus_newick.restype = vp
        L.phb_consensus_newick.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32]
        L.phb_cluster_component.argtypes = [vp, C.c_uint32, C.c_int, C.c_double, vp]
        L.phb_cluster_component_f64.argtypes = [vp, C.c_uint32, C.c_int, C.c_double, vp]
        L.phb_agglomerate.argtypes = [vp, C.c_uint32, C.c_int, C.c_double, vp, vp, vp, vp, u32p]
        L.phb_dendrogram_text.restype = vp
        L.phb_dendrogram_text.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, vp, vp, vp, C.c_uint32]
        L.phb_print_dist_mat.restype = vp
        L.phb_print_dist_mat.argtypes = [vp, C.c_uint32]
        L.phb_write_dist_mat.argtypes = [C.c_char_p, vp, C.c_uint32]
        L.phb_dist_bin_begin.argtypes = [C.c_char_p, C.c_uint32]
        L.phb_dist_bin_append.argtypes = [C.c_char_p, vp, C.c_uint64]
        L.phb_dist_bin_read.argtypes = [C.c_char_p, u32p, C.POINTER(vp)]
        L.phb_free_buf.argtypes = [vp]
        _host = L
    return _host


def rf_lib():
    """The CUDA library.  Raises (never falls back) when it is not built or cannot load."""
    global _rf
    if _rf is None:
        path = _build.RF_SO
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is not built: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "phybin_b200 has no CPU fallback for the RF distance matrix.")
        L = C.CDLL(path)
        vp = C.c_void_p
        st = C.POINTER(prf_stats)
        L.prf_create.restype = vp
        L.prf_create.argtypes = [C.c_int]
        L.prf_destroy.argtypes = [vp]
        L.prf_last_error.restype = C.c_char_p
        L.prf_last_error.argtypes = [vp]
        L.prf_version.restype = C.c_char_p
        L.prf_stream.restype = vp
        L.prf_stream.argtypes = [vp]
        L.prf_hashrf.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int32, vp, vp, st]
        L.prf_hashrf_load.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int, st]
        L.prf_hashrf_compute.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int32, C.c_int, vp, vp, st]
        L.prf_fetch.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
        L.prf_fetch_mask.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
        L.prf_fetch_rows.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
        L.prf_tolerant.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int32, vp, vp, st]
        L.prf_tolerant_load.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int, st]
        L.prf_tolerant_load_dups.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int, st]
        L.prf_tolerant_compute.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int32, vp, vp, st]
        L.prf_team_handle_bytes.restype = C.c_size_t
        L.prf_team_init.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, vp]
        L.prf_team_connect.argtypes = [vp, vp]
        L.prf_team_load.argtypes = [vp, vp, vp, C.c_uint32, st]
        L.prf_team_release.argtypes = [vp]
        L.prf_multi_create.restype = vp
        L.prf_multi_create.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.prf_multi_destroy.argtypes = [vp]
        L.prf_multi_last_error.restype = C.c_char_p
        L.prf_multi_last_error.argtypes = [vp]
        L.prf_multi_gpus.argtypes = [vp]
        L.prf_multi_ctx.restype = vp
        L.prf_multi_ctx.argtypes = [vp, C.c_int]
        L.prf_multi_hashrf.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int32, vp, vp, st]
        L.prf_allbips.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                  C.POINTER(prf_extract_info)]
        L.prf_hashrf_trees.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int32, vp, vp, st,
                                       C.POINTER(prf_extract_info)]
        L.prf_tolerant_trees.argtypes = L.prf_hashrf_trees.argtypes
        L.prf_allbips_result.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), u64p]
        L.prf_allbips_fetch.argtypes = [vp, vp, vp, vp]
        L.prf_fetch_submatrix.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, vp]
        L.prf_agglomerate.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, C.c_int, C.c_double, vp, vp, vp, vp,
                                      C.POINTER(prf_agglo_info)]
        L.prf_consensus.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int, vp, C.POINTER(prf_consensus_info)]
        L.prf_consensus_fetch.argtypes = [vp, vp, vp]
        L.prf_mask_components.argtypes = [vp, vp, C.c_uint32, vp, u32p]
        L.prf_mask_components_shard.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint32, vp, C.c_uint32, vp, u32p]
        L.prf_multi_components.argtypes = [vp, vp, u32p]
        L.prf_packed_index.restype = C.c_uint64
        L.prf_packed_index.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.prf_shard_align.restype = C.c_uint64
        L.prf_shard_range.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, u64p, u64p]
        L.prf_limits.argtypes = [u32p, u32p]
        L.prf_launch_count.restype = C.c_uint64
        L.prf_launch_count.argtypes = [vp]
        # test-only entry points (not in phybin_rf.h)
        L.prf_debug_small_tiles.argtypes = [vp, C.c_int]
        L.prf_debug_short_max.argtypes = [vp, C.c_uint32]
        L.prf_debug_rows_cfg.argtypes = [vp, C.c_int]
        L.prf_debug_no_flip.argtypes = [vp, C.c_int]
        L.prf_debug_row_plan.restype = C.c_longlong
        L.prf_debug_row_plan.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, vp, C.c_uint64]
        L.prf_debug_wide_slots.argtypes = [vp, C.c_int]
        L.prf_debug_agglo_ctas.argtypes = [vp, C.c_uint32]
        L.prf_debug_wordop_peak.argtypes = [vp, C.POINTER(C.c_double)]
        L.prf_debug_scan.argtypes = [vp, vp, C.c_uint64, vp, vp]
        L.prf_debug_sort.argtypes = [vp, vp, vp, C.c_uint32, C.c_int]
        _rf = L
    return _rf


# Every symbol include/phybin_rf.h declares (tests check the built library exports them all).
RF_SYMBOLS = [
    "prf_create", "prf_destroy", "prf_last_error", "prf_version", "prf_stream", "prf_hashrf",
    "prf_hashrf_load", "prf_hashrf_compute", "prf_multi_create", "prf_multi_destroy", "prf_multi_last_error", "prf_multi_gpus", "prf_multi_ctx", "prf_multi_hashrf", "prf_team_handle_bytes", "prf_team_init", "prf_team_connect", "prf_team_load", "prf_team_release", "prf_fetch", "prf_fetch_mask", "prf_fetch_rows",
    "prf_tolerant", "prf_tolerant_load", "prf_tolerant_load_dups", "prf_tolerant_compute", "prf_allbips", "prf_allbips_result", "prf_allbips_fetch", "prf_hashrf_trees", "prf_tolerant_trees", "prf_consensus", "prf_consensus_fetch", "prf_fetch_submatrix", "prf_agglomerate",
    "prf_mask_components", "prf_mask_components_shard", "prf_multi_components", "prf_packed_index",
    "prf_shard_align", "prf_shard_range", "prf_limits", "prf_launch_count",
]

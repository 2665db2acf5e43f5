/* phybin_host.h -- C interface of libphybin_host.so: the HOST side of the path, i.e. what stays
 * in the Haskell program in a real integration (Newick parsing, the global label table, the
 * HashRF leaf-count filter, bipartition extraction `allBips`, printDistMat) re-stated in C++
 * because this image has no GHC.  It produces exactly the arrays include/phybin_rf.h consumes.
 * It contains no CUDA and never touches the oracle.
 *
 * Reference interfaces mirrored (all under /root/reference/src/Bio/Phylogeny/PhyBin/):
 *   parseNewicks      Parser.hs:57-75      (label order: last file first, children right-to-left)
 *   numLeaves         CoreTypes.hs:289-291
 *   allBips/normBip   RFDistance.hs:207-248
 *   treeLabels        RFDistance.hs:200-203
 *   printDistMat      RFDistance.hs:352-361
 *   HashRF filter     ../PhyBin.hs:187-194
 */
#ifndef PHYBIN_HOST_H
#define PHYBIN_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PHB_API __attribute__((visibility("default")))
#else
#define PHB_API
#endif

typedef struct phb_forest phb_forest;

/* Synthetic tree families (BASELINE.md section 3 / SURVEY.md 8d). */
typedef enum phb_family {
  PHB_RANDOM_TOPOLOGY = 0, /* uniform unrooted binary tree by random-edge insertion */
  PHB_NNI_FAMILY = 1       /* one random base tree + r ~ U{0..nni_max} random NNI moves per tree */
} phb_family;

/* Parses `nfiles` Newick texts (command-line order); name_prefix >= 0 applies `take n` to leaf
 * names (-p, Main.hs:290).  Never returns NULL; check phb_error(). */
PHB_API phb_forest* phb_parse(int nfiles, const char* const* texts, int name_prefix);
/* T trees on taxa t0..t{n-1}; splitmix64 stream seeded with `seed`; each taxon is dropped from a
 * tree independently with probability p_missing (at least 4 are kept).  Labels are assigned as
 * parseNewicks would assign them to the emitted Newick text. */
PHB_API phb_forest* phb_synth(uint32_t T, uint32_t n, uint64_t seed, int family, uint32_t nni_max,
                              double p_missing);
PHB_API void phb_free(phb_forest* f);
PHB_API const char* phb_error(const phb_forest* f); /* NULL when fine */

PHB_API uint32_t phb_num_trees(const phb_forest* f);
PHB_API uint32_t phb_num_labels(const phb_forest* f);
PHB_API const char* phb_label_name(const phb_forest* f, uint32_t label);
PHB_API uint32_t phb_tree_num_leaves(const phb_forest* f, uint32_t t);
/* Trees parsed from the i-th text handed to phb_parse (as parsed: filters do not update it). */
PHB_API uint32_t phb_file_num_trees(const phb_forest* f, uint32_t i);
/* Drops trees whose numLeaves != expected (PhyBin.hs:187-194); returns the number kept. */
PHB_API uint32_t phb_filter_num_leaves(phb_forest* f, uint32_t expected);
/* Keeps only trees [t0, t1). */
PHB_API uint32_t phb_slice(phb_forest* f, uint32_t t0, uint32_t t1);
/* 1 when some tree repeats a leaf label. */
PHB_API int phb_has_duplicate_leaves(const phb_forest* f);

/* Newick text of trees [t0, t1), one per line; free with phb_free_buf. */
PHB_API char* phb_to_newick(const phb_forest* f, uint32_t t0, uint32_t t1);

/* Words per bit set needed for this forest: ceil(max(#labels, max numLeaves) / 64). */
PHB_API uint32_t phb_words(const phb_forest* f);

/* allBips of every tree, canonicalised exactly as normBip does and bit-packed (hashRF's input).
 * *bips: nnz*W words, *tree_off: T+1 offsets; both malloc'd, free with phb_free_buf.
 * threads <= 0: use all hardware threads. */
PHB_API int phb_allbips(const phb_forest* f, uint32_t W, int threads, uint64_t** bips,
                        uint64_t** tree_off, uint64_t* nnz);

/* Tolerant-mode inputs: un-normalised leaf sets of the non-root interior nodes + per-tree taxon
 * sets.  *clades: nnz*W, *tree_off: T+1, *leafsets: T*W. */
PHB_API int phb_raw_clades(const phb_forest* f, uint32_t W, int threads, uint64_t** clades,
                           uint64_t** tree_off, uint64_t** leafsets, uint64_t* nnz);

/* Tolerant-mode inputs for forests in which a tree may REPEAT a leaf label (prf_tolerant_load_dups, phybin_rf.h).
 * The reference counts such leaves twice in numLeaves (CoreTypes.hs:289-291) -- normBip's size -- while its
 * label sets stay sets, so the device also needs
 *   *outsets    nnz*W: per clade, the label set of the leaves OUTSIDE that node (decides whether pruneTreeLeaves,
 *               PreProcessor.hs:21-32, leaves the node below the pruned root)
 *   *dup_off    T+1 offsets into dup_label / dup_extra: tree t's labels that occur more than once and how many
 *               extra times.
 * Clades are every non-root interior node in pre-order (not made unique).  All arrays malloc'd, free with
 * phb_free_buf. */
PHB_API int phb_raw_clades_ex(const phb_forest* f, uint32_t W, uint64_t** clades, uint64_t** outsets,
                              uint64_t** tree_off, uint64_t** leafsets, uint64_t** dup_off, uint32_t** dup_label,
                              uint32_t** dup_extra, uint64_t* nnz);

/* The trees as the flat pre-order arrays prf_allbips (phybin_rf.h) takes: BORROWED pointers into the
 * forest, valid until it is modified or freed.  node_label[v] >= 0: leaf label, -1: interior node;
 * node_parent[v]: index of the parent (undefined for a tree's first node, its root); the nodes of tree
 * t are tree_begin[t] .. tree_begin[t+1]-1; num_leaves[t] = numLeaves.  Any out pointer may be NULL. */
PHB_API int phb_tree_arrays(const phb_forest* f, const int32_t** node_label, const uint32_t** node_parent,
                            const uint64_t** tree_begin, const uint32_t** num_leaves, uint64_t* n_nodes);

/* bipsToTree num_taxa bips (RFDistance.hs:410-445) followed by displayStrippedTree (CoreTypes.hs:127-132) on one
 * line: the tree whose bipartitions are the k given W-word sets (e.g. a consensus set from prf_consensus), with
 * the forest's label names, as "(a, (b, c), d);".  Same subtree order as the reference: bipartitions by increasing
 * size (ties in IntSet order), matching subtrees glommed in list order, the new node put in front.  NULL when a
 * set has no subtree inside it or names a label the forest does not have; free with phb_free_buf. */
PHB_API char* phb_consensus_newick(const phb_forest* f, const uint64_t* keys, uint64_t k, uint32_t W,
                                   uint32_t num_taxa);

/* Flat clusters of ONE connected component of { d <= thresh } under complete linkage or UPGMA -- the part of
 * C.dendrogram + sliceDendro (PhyBin.hs:358,415-422) that is left once prf_mask_components has split the trees
 * (single linkage needs nothing more: its flat clusters are the components).
 *   sub      packed upper triangle (m*(m-1)/2) of the component's m x m distance matrix (prf_fetch_submatrix)
 *   linkage  0 single, 1 complete, 2 UPGMA (C.Linkage; Main.hs:96-98)
 *   labels   out, m entries: smallest local index of every tree's flat cluster
 * Agglomeration as hierarchical-clustering-0.4.6's distance-matrix algorithm does it (stack.yaml:9; source NOT
 * under /root/reference, restated from the published package -- tie order unpinned): repeatedly merge the pair
 * of clusters at minimal distance, first in (lower key, higher key) lexicographic order among ties; the merged
 * cluster keeps the lower key; distances follow Lance-Williams (min / max / size-weighted mean).  Merging stops
 * at the first distance > thresh, which for these monotone linkages is exactly sliceDendro's cut.
 * Returns 0, -1 on bad arguments, -2 when m is too large for a dense m x m matrix of doubles. */
PHB_API int phb_cluster_component(const uint16_t* sub, uint32_t m, int linkage, double thresh, uint32_t* labels);
/* Same on doubles (tests against scipy use tie-free real distances). */
PHB_API int phb_cluster_component_f64(const double* sub, uint32_t m, int linkage, double thresh, uint32_t* labels);

/* The whole agglomeration with its merge list (the host twin of prf_agglomerate, phybin_rf.h; same rule, same
 * IEEE-double arithmetic): step k merged the clusters with smallest members merge_a[k] < merge_b[k] at distance
 * merge_d[k]; *n_merges steps were taken before the smallest distance exceeded thresh (m - 1 with
 * thresh = INFINITY).  labels / merge_* (m - 1 entries each) / n_merges may be NULL. */
PHB_API int phb_agglomerate(const uint16_t* sub, uint32_t m, int linkage, double thresh, uint32_t* labels,
                            uint32_t* merge_a, uint32_t* merge_b, double* merge_d, uint32_t* n_merges);

/* dendrogram.txt: `show (fmap treename dendro)` (PhyBin.hs:238-243) for a FULL merge list (n_merges == m - 1) --
 * the derived Show of hierarchical-clustering's `Dendrogram String`,
 *   Leaf "name"   |   Branch <show Double> (<left>) (<right>)
 * with the cluster that kept its key (merge_a) as the left child.  names: m tree names.  NULL on an incomplete
 * or malformed merge list; free with phb_free_buf. */
PHB_API char* phb_dendrogram_text(const char* const* names, uint32_t m, const uint32_t* merge_a,
                                  const uint32_t* merge_b, const double* merge_d, uint32_t n_merges);

/* printDistMat text of a packed uint16 upper triangle; free with phb_free_buf. */
PHB_API char* phb_print_dist_mat(const uint16_t* upper, uint32_t T);
/* Streams the same text to a file; returns 0 on success. */
PHB_API int phb_write_dist_mat(const char* path, const uint16_t* upper, uint32_t T);

/* Binary form of the matrix for sizes where the text dump is unusable.  Layout, little-endian:
 *   bytes 0-7 "PHYBINRF", uint32 version = 1, uint32 T, uint64 P = T(T-1)/2, uint32 bits = 16, uint32 layout = 0
 *   (packed row-major upper triangle, phybin_rf.h), then P uint16 entries.
 * begin truncates the file and writes the header; append adds entries in packed order (so a caller can stream
 * prf_fetch chunks without ever holding the matrix); read returns a malloc'd copy (free with phb_free_buf).
 * All return 0 on success. */
PHB_API int phb_dist_bin_begin(const char* path, uint32_t T);
PHB_API int phb_dist_bin_append(const char* path, const uint16_t* entries, uint64_t n);
PHB_API int phb_dist_bin_read(const char* path, uint32_t* T, uint16_t** upper);

PHB_API void phb_free_buf(void* p);

#ifdef __cplusplus
}
#endif
#endif /* PHYBIN_HOST_H */
